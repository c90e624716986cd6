// Host-side constant tables of the encode path: ITU-T T.81 Annex K Huffman
// specifications (the reference spells the resulting codes out as strings in
// src/huffman.hpp:9,26,43,250), canonical code generation (T.81 Annex C), the
// quantisation constants of the fused kernel and the JFIF header writer (the
// reference has no file writer; segment order per SURVEY.md section 8c).
#include <math.h>
#include <string.h>

#include "jb_internal.h"

namespace jb {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

namespace {

struct HuffSpec {
    uint8_t bits[16];
    const uint8_t* vals;
    int nvals;
};

const uint8_t kDcVals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
// K.5: run/size symbols ordered by code length
const uint8_t kAcLumVals[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
// K.6
const uint8_t kAcChrVals[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

// order of the DHT segments: DC luma (00), AC luma (10), DC chroma (01), AC chroma (11)
const HuffSpec kSpecs[4] = {
    {{0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0}, kDcVals, 12},           // K.3
    {{0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d}, kAcLumVals, 162},    // K.5
    {{0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0}, kDcVals, 12},           // K.4
    {{0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77}, kAcChrVals, 162}};   // K.6
const uint8_t kTcTh[4] = {0x00, 0x10, 0x01, 0x11};

void assign_codes(const HuffSpec& sp, uint32_t* table /* indexed by symbol */) {
    uint32_t code = 0;
    int k = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < sp.bits[len - 1]; ++i) table[sp.vals[k++]] = (code++ << 5) | (uint32_t)len;
        code <<= 1;
    }
}

}  // namespace

// run/size code and value bits of small coefficients in one entry (utils.cpp:623-653, 683-691)
static void build_small(HuffDev* out) {
    for (int t = 0; t < 2; ++t)
        for (int run = 0; run < 16; ++run)
            for (int v = -15; v <= 15; ++v) {
                if (v == 0) continue;
                int a = v < 0 ? -v : v, cat = 0;
                while (a >> cat) ++cat;
                uint32_t vb = (uint32_t)(v < 0 ? v + (1 << cat) - 1 : v);
                uint32_t e = out->ac[t][(run << 4) | cat];
                out->small[t][(run << 5) | (v & 31)] = ((((e >> 5) << cat) | vb) << 5) | ((e & 31u) + (uint32_t)cat);
            }
}

// (code << 5) | length, the form the builders above work in, to the device form: code left-aligned | length
static void left_align(HuffDev* out) {
    uint32_t* tabs[3] = {&out->ac[0][0], &out->dc[0][0], &out->small[0][0]};
    const int n[3] = {512, 32, 1024};
    for (int t = 0; t < 3; ++t)
        for (int i = 0; i < n[t]; ++i) {
            const uint32_t e = tabs[t][i], len = e & 31u;
            tabs[t][i] = len ? ((e >> 5) << (32 - len)) | len : 0u;
        }
}

void annex_k_specs(HuffSpecs* sp) {
    memset(sp, 0, sizeof(*sp));
    for (int t = 0; t < 4; ++t) {
        memcpy(sp->bits[t], kSpecs[t].bits, 16);
        memcpy(sp->vals[t], kSpecs[t].vals, (size_t)kSpecs[t].nvals);
        sp->n[t] = kSpecs[t].nvals;
    }
}

void build_huff_from_specs(const HuffSpecs& sp, HuffDev* out) {
    memset(out, 0, sizeof(*out));
    uint32_t* tables[4] = {out->dc[0], out->ac[0], out->dc[1], out->ac[1]};
    for (int t = 0; t < 4; ++t) {
        HuffSpec one;
        memcpy(one.bits, sp.bits[t], 16);
        one.vals = sp.vals[t];
        one.nvals = sp.n[t];
        uint32_t full[256] = {0};
        assign_codes(one, full);
        memcpy(tables[t], full, (t & 1 ? 256 : 16) * sizeof(uint32_t));
    }
    build_small(out);
    left_align(out);
}

// Optimal BITS / HUFFVAL for 256 symbol counts: T.81 Annex K.2 (Figures K.1-K.4) in the form libjpeg's
// jpeg_gen_optimal_table gives it -- a reserved 257th symbol of count 1 keeps the all-ones code free, the two
// least frequent entries merge (the larger symbol on ties), code lengths above 16 are folded back.
void optimal_spec(const uint64_t counts[256], uint8_t bits_out[16], uint8_t vals_out[256], int* n_out) {
    const int kMax = 32;
    uint64_t base[256], w[257];
    int len[257], next[257], hist[kMax + 1];
    for (int i = 0; i < 256; ++i) base[i] = counts[i];
    for (bool again = true; again;) {
        for (int i = 0; i < 256; ++i) w[i] = base[i];
        w[256] = 1;
        for (int i = 0; i < 257; ++i) { len[i] = 0; next[i] = -1; }
        auto least = [&](int skip) {
            int best = -1;
            for (int i = 0; i <= 256; ++i)
                if (w[i] && i != skip && (best < 0 || w[i] <= w[best])) best = i;
            return best;
        };
        for (;;) {
            int a = least(-1), b = a < 0 ? -1 : least(a);
            if (b < 0) break;
            w[a] += w[b];
            w[b] = 0;
            // every symbol of both chains moves one level down; chain b is appended to chain a
            for (int i = a;; i = next[i]) {
                ++len[i];
                if (next[i] < 0) { next[i] = b; break; }
            }
            for (int i = b; i >= 0; i = next[i]) ++len[i];
        }
        again = false;
        for (int i = 0; i <= 256; ++i)
            if (len[i] > kMax) again = true;
        // a tree deeper than 32 (libjpeg stops with an error there; it takes millions of symbols in a
        // Fibonacci-like distribution): halve every count, rounding up, and build it again
        if (again)
            for (int i = 0; i < 256; ++i)
                if (base[i]) base[i] = (base[i] + 1) / 2;
    }
    for (int i = 0; i <= kMax; ++i) hist[i] = 0;
    for (int i = 0; i <= 256; ++i)
        if (len[i]) ++hist[len[i]];
    int l = kMax;
    for (; l > 16; --l)
        while (hist[l] > 0) {  // Figure K.3
            int j = l - 2;
            while (hist[j] == 0) --j;
            hist[l] -= 2;
            ++hist[l - 1];
            hist[j + 1] += 2;
            --hist[j];
        }
    while (hist[l] == 0) --l;
    --hist[l];  // the reserved symbol
    for (int i = 1; i <= 16; ++i) bits_out[i - 1] = (uint8_t)hist[i];
    int n = 0;
    memset(vals_out, 0, 256);
    for (int i = 1; i <= kMax; ++i)
        for (int sym = 0; sym < 256; ++sym)
            if (len[sym] == i) vals_out[n++] = (uint8_t)sym;
    *n_out = n;
}

void optimal_huff_specs(const uint64_t counts[4][256], HuffSpecs* sp) {
    memset(sp, 0, sizeof(*sp));
    for (int t = 0; t < 4; ++t) optimal_spec(counts[t], sp->bits[t], sp->vals[t], &sp->n[t]);
}

void build_huff(bool typo, HuffDev* out) {
    memset(out, 0, sizeof(*out));
    assign_codes(kSpecs[0], out->dc[0]);
    assign_codes(kSpecs[1], out->ac[0]);
    assign_codes(kSpecs[2], out->dc[1]);
    assign_codes(kSpecs[3], out->ac[1]);
    if (typo) {
        // huffman.hpp:92-98: the strings for luma AC 3/4 .. 3/A have 17 characters,
        // the 16-bit Annex-K code with an extra leading '1' (SURVEY Q2).
        for (int cat = 4; cat <= 10; ++cat) {
            uint32_t e = out->ac[0][(3 << 4) | cat];
            uint32_t code = (e >> 5) | (1u << 16);
            out->ac[0][(3 << 4) | cat] = (code << 5) | 17u;
        }
    }
    build_small(out);
    left_align(out);
}

// ---- worst-case error of the binary32 AAN transform -------------------------------
// Forward error analysis of fdct8 (jb_math.h) applied to rows, then columns, of
// integer samples |x| <= 128: every node carries (max |value|, bound on |computed -
// exact|).  Additions of integers below 2^24 are exact; every other operation adds
// the unit round-off u = 2^-24 times its result magnitude, and every irrational
// constant its representation error.  The result bounds |a - a_exact| for each of
// the 64 AAN-scaled outputs; tests/test_math_host.py checks measured errors against it.
namespace {
struct Bnd {
    double m, e;
    bool isint;
};
const double kU = 5.9604644775390625e-08;  // 2^-24
Bnd b_add(Bnd a, Bnd b) {
    Bnd r;
    r.m = a.m + b.m;
    r.isint = a.isint && b.isint && r.m < 16777216.0;
    r.e = a.e + b.e + (r.isint ? 0.0 : kU * (r.m + a.e + b.e));
    return r;
}
Bnd b_fma(Bnd w, double c, Bnd x) {
    Bnd r;
    c = fabs(c);
    r.m = w.m * c + x.m;
    r.isint = false;
    double e = w.e * c + w.m * c * kU + x.e;
    r.e = e + kU * (r.m + e);
    return r;
}
void b_fdct8(Bnd* d) {
    const Bnd zero{0, 0, true};
    Bnd s07 = b_add(d[0], d[7]), s16 = b_add(d[1], d[6]), s25 = b_add(d[2], d[5]), s34 = b_add(d[3], d[4]);
    Bnd e0 = b_add(s07, s34), e1 = b_add(s16, s25);  // differences have the same bounds as sums
    Bnd o[8];
    o[0] = o[4] = b_add(e0, e1);
    Bnd w = b_add(e1, e0);
    o[2] = o[6] = b_fma(w, JB_C4, e0);
    Bnd o0 = b_add(s34, s25), o1 = b_add(s25, s16), o2 = b_add(s16, s07);
    Bnd z5 = b_fma(b_add(o0, o2), JB_C6, zero);
    Bnd z2 = b_fma(o0, JB_Q, z5), z4 = b_fma(o2, JB_R, z5), z11 = b_fma(o1, JB_C4, s07);
    o[5] = o[3] = b_add(z11, z2);
    o[1] = o[7] = b_add(z11, z4);
    for (int i = 0; i < 8; ++i) d[i] = o[i];
}
}  // namespace

void aan_error_bound(double err[64], double amax[64]) {
    Bnd row[8];
    for (int i = 0; i < 8; ++i) row[i] = Bnd{128.0, 0.0, true};
    b_fdct8(row);
    for (int u = 0; u < 8; ++u) {
        Bnd col[8];
        for (int i = 0; i < 8; ++i) col[i] = row[u];
        b_fdct8(col);
        for (int v = 0; v < 8; ++v) {
            err[v * 8 + u] = col[v].e;
            amax[v * 8 + u] = col[v].m;
        }
    }
}

void build_quant_const(const uint32_t ql[64], const uint32_t qc[64], QuantConst* out) {
    double err[64], amax[64];
    aan_error_bound(err, amax);
    for (int t = 0; t < 2; ++t) {
        const uint32_t* q = t ? qc : ql;
        for (int v = 0; v < 8; ++v)
            for (int u = 0; u < 8; ++u) {
                int n = v * 8 + u;
                double alpha = (u == 0 ? M_SQRT1_2 : 1.0) * (v == 0 ? M_SQRT1_2 : 1.0);
                double mul = alpha / (4.0 * (double)q[n] * aan_scale(u) * aan_scale(v));
                out->mul[t][n] = (float)mul;
                // error of a*mul: transform error, the multiplier's own rounding, and the two roundings
                // inside quantize_bits; 25 % margin on top of the worst case.
                double delta = 1.25 * ((err[n] + 2.0 * kU * amax[n]) * mul) + 1e-6;
                float band = (float)(0.5 - delta);
                out->band[t][n] = nextafterf(band, 0.0f);  // float(0.5 - delta) may have rounded up
            }
        // DC in integers: S / (8 q) with exact ties going towards zero.  Check the rule against the
        // reference's binary64 expression (utils.cpp:336, 460) for every possible sample sum S.
        uint32_t D = 8 * q[0];
        out->dc_d[t] = D;
        out->dc_m[t] = (uint32_t)((0x100000000ull + 2 * D - 1) / (2 * D));
    }
    double a0 = 1.0 / sqrt(2);
    volatile double scale00 = (a0 * a0 / 4.0);
    bool ok = true;
    for (int t = 0; t < 2 && ok; ++t) {
        uint32_t D = out->dc_d[t];
        double q = (double)(t ? qc : ql)[0];
        for (int S = -8192; S <= 8192 && ok; ++S) {
            volatile double F = (double)S * scale00;
            volatile double quo = F / q;
            int want = (int)round(quo);
            uint32_t A = (uint32_t)(S < 0 ? -S : S);
            uint32_t m = (uint32_t)(((uint64_t)(2 * A + D - 1) * out->dc_m[t]) >> 32);
            int got = S < 0 ? -(int)m : (int)m;
            ok = got == want;
        }
    }
    out->dc_exact = ok ? 1u : 0u;
}

// Y tie table (jb_math.h): bit (r<<8|g) set when the reference's binary64 expression
// (utils.cpp:107) lands below the exact integer value of 0.299r + 0.587g + 0.114b.
void build_ydown(uint32_t ydown[2048]) {
    memset(ydown, 0, 2048 * sizeof(uint32_t));
    for (uint32_t r = 0; r < 256; ++r)
        for (uint32_t g = 0; g < 256; ++g)
            for (uint32_t b = 0; b < 256; ++b) {
                uint32_t s = 299 * r + 587 * g + 114 * b;
                if (s % 1000) continue;
                volatile double y = 0.299 * r + 0.587 * g + 0.114 * b;
                if ((uint32_t)(uint8_t)y != s / 1000) ydown[(r << 8 | g) >> 5] |= 1u << ((r << 8 | g) & 31);
            }
}

// The cosine and scale factors of the binary64 replay, formed exactly as the reference forms
// them (utils.cpp:330-332 and 317-318, 336) with the same libm.
void build_dct_tables(double costab[64], double scale[64]) {
    for (size_t u = 0; u < 8; ++u)
        for (size_t x = 0; x < 8; ++x) costab[u * 8 + x] = cos((2 * x + 1) * u * M_PI / 16.0);
    for (size_t u = 0; u < 8; ++u)
        for (size_t v = 0; v < 8; ++v) {
            double alphaU = (u == 0) ? 1.0 / sqrt(2) : 1.0;
            double alphaV = (v == 0) ? 1.0 / sqrt(2) : 1.0;
            scale[u * 8 + v] = (alphaU * alphaV / 4.0);
        }
}

static void put16(uint8_t* p, unsigned v) {
    p[0] = (uint8_t)(v >> 8);
    p[1] = (uint8_t)v;
}

size_t build_header(const jb_params* p, size_t W, size_t H, uint8_t* h, const HuffSpecs* custom) {
    HuffSpecs std_specs;
    if (!custom) {
        annex_k_specs(&std_specs);
        custom = &std_specs;
    }
    size_t n = 0;
    static const uint8_t soi_app0[] = {0xFF, 0xD8, 0xFF, 0xE0, 0x00, 0x10, 'J',  'F',  'I',  'F',
                                       0x00, 0x01, 0x01, 0x00, 0x00, 0x01, 0x00, 0x01, 0x00, 0x00};
    memcpy(h, soi_app0, sizeof(soi_app0));
    n += sizeof(soi_app0);
    for (int t = 0; t < 2; ++t) {  // DQT, 8-bit entries in zigzag order
        h[n++] = 0xFF; h[n++] = 0xDB; h[n++] = 0x00; h[n++] = 0x43; h[n++] = (uint8_t)t;
        for (int k = 0; k < 64; ++k) h[n++] = (uint8_t)(t ? p->qchrom : p->qlum)[kZigzag[k]];
    }
    unsigned dw = W > 65535 ? 65535u : (unsigned)W, dh = H > 65535 ? 65535u : (unsigned)H;
    h[n++] = 0xFF; h[n++] = 0xC0; h[n++] = 0x00; h[n++] = 0x11; h[n++] = 0x08;  // SOF0
    put16(h + n, dh); n += 2;
    put16(h + n, dw); n += 2;
    h[n++] = 3;
    h[n++] = 1; h[n++] = (uint8_t)(p->subsampling == JB_SUB_420 ? 0x22 : 0x11); h[n++] = 0;
    h[n++] = 2; h[n++] = 0x11; h[n++] = 1;
    h[n++] = 3; h[n++] = 0x11; h[n++] = 1;
    for (int t = 0; t < 4; ++t) {  // DHT x4
        h[n++] = 0xFF; h[n++] = 0xC4;
        put16(h + n, (unsigned)(2 + 1 + 16 + custom->n[t])); n += 2;
        h[n++] = kTcTh[t];
        memcpy(h + n, custom->bits[t], 16); n += 16;
        memcpy(h + n, custom->vals[t], (size_t)custom->n[t]); n += (size_t)custom->n[t];
    }
    if (p->restart_interval > 0) {  // DRI
        h[n++] = 0xFF; h[n++] = 0xDD; h[n++] = 0x00; h[n++] = 0x04;
        put16(h + n, (unsigned)p->restart_interval); n += 2;
    }
    static const uint8_t sos[] = {0xFF, 0xDA, 0x00, 0x0C, 0x03, 0x01, 0x00, 0x02, 0x11, 0x03, 0x11, 0x00, 0x3F, 0x00};
    memcpy(h + n, sos, sizeof(sos));
    n += sizeof(sos);
    return n;
}

// ---- tensor-core variant: the whole block transform as one 64x64 contraction -----------
// t[n] = sum_k x[k] * W[n][k], k = y*8+x (sample), n = zigzag position, with
//   W[n][k] = alpha(u)alpha(v)/4 * cos((2x+1)u pi/16) cos((2y+1)v pi/16) / q[v][u],  (v,u) = zigzag(n)
// i.e. FDCT, quantiser scale and zigzag permutation in a single matrix.  W is scaled by 2^10
// (undone for free by the FMA that rounds) and split into two fp16 matrices, hi + lo = 22
// significand bits; the 8-bit samples are exact in fp16.  Layout = UMMA B operand: 64 rows
// (n) x 128 bytes (k), K-major, 128-byte swizzle (16-byte chunk c of row n at
// n*128 + ((c ^ (n&7)) << 4)).  out = [table][split] x 8192 bytes.
// tband[t][n] = 0.5 - (derived error bound of coefficient n), see build_tc_matrices.
static uint16_t to_fp16(double x, double* back) {
    // round to nearest even into IEEE binary16 (values here are far from overflow)
    if (x == 0.0) { *back = 0.0; return 0; }
    int sign = x < 0; double ax = fabs(x);
    int e; double m = frexp(ax, &e);          // ax = m * 2^e, m in [0.5,1)
    int E = e - 1;                             // ax = (2m) * 2^E, 2m in [1,2)
    uint16_t h;
    if (E < -14) {                             // subnormal: units of 2^-24
        double q = nearbyint(ax * 16777216.0);
        h = (uint16_t)q;
        *back = q / 16777216.0;
    } else {
        double q = nearbyint((2.0 * m - 1.0) * 1024.0);  // 10 fraction bits
        if (q == 1024.0) { q = 0; ++E; }
        h = (uint16_t)(((E + 15) << 10) | (int)q);
        *back = ldexp(1.0 + q / 1024.0, E);
    }
    if (sign) { h |= 0x8000; *back = -*back; }
    return h;
}

// Q1 (utils.cpp:314-347): performDCTBlock writes every output into the block it is still reading (outputs in
// the order u outer / v inner, F(u,v) stored at row v, column u).  The result is still a linear function of the 64
// samples: map[nat][k] = output `nat` for the unit block e_k, obtained by running that very loop in binary64.
static void inplace_dct_map(double map[64][64]) {
    const double pi = 3.14159265358979323846;
    double c[8][8];
    for (int u = 0; u < 8; ++u)
        for (int x = 0; x < 8; ++x) c[u][x] = cos((2 * x + 1) * u * pi / 16.0);
    for (int k = 0; k < 64; ++k) {
        double blk[64] = {0};
        blk[k] = 1.0;
        for (int u = 0; u < 8; ++u)
            for (int v = 0; v < 8; ++v) {
                double s = 0.0;
                for (int y = 0; y < 8; ++y)
                    for (int x = 0; x < 8; ++x) s += blk[y * 8 + x] * c[u][x] * c[v][y];
                blk[v * 8 + u] = s * ((u == 0 ? M_SQRT1_2 : 1.0) * (v == 0 ? M_SQRT1_2 : 1.0) / 4.0);
            }
        for (int n = 0; n < 64; ++n) map[n][k] = blk[n];
    }
}

void build_tc_matrices(const uint32_t ql[64], const uint32_t qc[64], double step_ulps, int repl_chroma, int inplace_dct,
                       uint8_t* out, float tband[2][64]) {
    // W[n][k] = (2-D DCT basis) x (1 / quantiser) x 2^10, row n in zigzag order, as TWO fp16 matrices in FIXED POINT:
    //   hi[n][k] = Qhi(n) * round(W / Qhi),      Qhi(n) = 2^(E-10),  2^E >= max_k |W[n][k]|   (|integer| <= 2^10)
    //   lo[n][k] = q0(n)  * round((W - hi) / q0), q0(n)  = Qhi / 2^11                          (|integer| <= 2^10)
    // Both are exact in binary16 (11 significand bits; the scan of tests/tools/tc_model_scan.py shows that the tensor
    // core also takes binary16 subnormals exactly).  The kernel issues the four K-chunks of `hi` first, then the four of
    // `lo`, into one fp32 accumulator.  What that costs in accuracy follows from two properties of the MMA datapath,
    // both pinned on the B200 by tests/test_gpu_tc_model.py with exact integer references (profiles/r02_tc_numerics.md):
    //  (P1) bits at or above the ulp of the largest addend of an MMA step are never discarded.  In the hi phase every
    //       product (8-bit sample x hi) is a multiple of Qhi and every partial sum is below 128 * 64 * 2^10 Qhi =
    //       2^23 Qhi < 2^24 Qhi, so ulp(largest) <= Qhi and the hi phase is EXACT.
    //  (P2) one step (accumulator + 16 products) deviates from the exact sum by less than `step_ulps` ulps of the
    //       largest magnitude involved: the addends are aligned to the largest exponent with 3 guard bits and
    //       truncated (16 * 2^-3 ulp), then the sum is truncated to binary32 (1 ulp) -- measured maximum 2.95,
    //       JB_TC_STEP_ULPS = 4.  Every magnitude is at most B(n) = 128 * sum_k (|hi| + |lo|), whose ulp is U(n).
    // Hence |accumulator - 2^10 F/q| <= 4 steps * step_ulps * U(n) + 128 * sum_k |W - hi - lo| (the representation
    // residual, summed exactly here), and a coefficient is flagged for the binary64 replay when its quotient lies
    // within that distance (+ 1e-6 for the two roundings of the epilogue) of a rounding tie.
    const double pi = 3.14159265358979323846;
    memset(out, 0, 32768);
    struct Q1Map {
        double m[64][64];
        Q1Map() { inplace_dct_map(m); }
    };
    static const Q1Map q1;  // built once (thread-safe static initialisation)
    const double (*q1_map)[64] = q1.m;
    for (int t = 0; t < 2; ++t) {
        const uint32_t* q = t ? qc : ql;
        const bool cells = t == 1 && repl_chroma;  // K = 16: one column per 2x2 cell, the sum of its four entries
        for (int n = 0; n < 64; ++n) {
            int nat = kZigzag[n], v = nat >> 3, u = nat & 7;
            double alpha = (u == 0 ? M_SQRT1_2 : 1.0) * (v == 0 ? M_SQRT1_2 : 1.0) / 4.0;
            double w[64];
            for (int k = 0; k < 64; ++k) {
                int y = k >> 3, x = k & 7;
                w[k] = alpha * cos((2 * x + 1) * u * pi / 16.0) * cos((2 * y + 1) * v * pi / 16.0) / (double)q[nat];
                if (inplace_dct) w[k] = q1_map[nat][k] / (double)q[nat];
            }
            const int nk = cells ? 16 : 64;
            double W[64], m = 0;
            for (int k = 0; k < nk; ++k) {
                double wk = w[k];
                if (cells) {
                    const int i = k >> 2, j = k & 3, k0 = (2 * i) * 8 + 2 * j;
                    wk = (w[k0] + w[k0 + 1]) + (w[k0 + 8] + w[k0 + 9]);
                }
                W[k] = wk * JB_TC_W_SCALE;
                m = fmax(m, fabs(W[k]));
            }
            if (m == 0.0) {  // an all-zero row (K = 16 cells: the basis functions with u = 4 or v = 4 cancel inside every 2x2
                             // cell): the matrices stay zero, the accumulator is exactly 0, nothing to replay
                tband[t][n] = nextafterf(0.5f, 0.0f);
                continue;
            }
            int E = 0;
            frexp(m, &E);          // m = f * 2^E with f in [0.5, 1): every |W| is below 2^E
            if (E < -3) E = -3;    // q0 = 2^(E-21) stays a binary16 (subnormal) quantum, 2^-24
            bool ok = m > 0 && E <= 15;
            const double Qhi = ldexp(1.0, E - 10), q0 = ldexp(1.0, E - 21);
            double sum_abs = 0, resid = 0;
            for (int k = 0; k < nk; ++k) {
                const double hi = nearbyint(W[k] / Qhi) * Qhi, lo = nearbyint((W[k] - hi) / q0) * q0;
                double b0 = 0, b1 = 0;
                uint16_t h0 = ok ? to_fp16(hi, &b0) : 0, h1 = ok ? to_fp16(lo, &b1) : 0;
                ok = ok && b0 == hi && b1 == lo && fabs(hi) <= 1024.0 * Qhi && fabs(lo) <= 1024.0 * q0;
                sum_abs += fabs(hi) + fabs(lo);
                resid += fabs(W[k] - hi - lo);
                size_t off = (size_t)n * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
                memcpy(out + ((size_t)t * 2 + 0) * 8192 + off, &h0, 2);
                memcpy(out + ((size_t)t * 2 + 1) * 8192 + off, &h1, 2);
            }
            if (!ok) {  // not representable (never with 8-bit quantisers and the true DCT): every coefficient of the row is replayed
                tband[t][n] = -1.0f;
                continue;
            }
            // (a row whose entries all round to zero -- K = 16 cells: the basis functions with u = 4 or v = 4 cancel inside
            // every 2x2 cell up to ~1e-16 -- leaves the accumulator exactly 0: no rounding step, only the residual.
            // ilogb(0) - 23 wrapped around, U became infinite and every such coefficient was replayed: 249 M per
            // 256 frames in the reference's own mode, 12 ms + 46 ms instead of 1.1 ms + 0.06 ms.)
            const double B = 128.0 * sum_abs, U = B > 0 ? ldexp(1.0, ilogb(B) - 23) : 0.0;
            const double err = 4.0 * step_ulps * U + 128.0 * resid;
            float band = (float)(0.5 - err / JB_TC_W_SCALE - 1e-6);
            tband[t][n] = nextafterf(band, 0.0f);
        }
    }
}

}  // namespace jb

// ---- host-only entry points of the C ABI (no device needed) ---------------------------
extern "C" int jb_optimal_huffman_spec(const uint64_t counts[256], uint8_t bits[16], uint8_t vals[256], int* n_vals) {
    if (!counts || !bits || !vals) return JB_E_INVALID;
    int n = 0;
    jb::optimal_spec(counts, bits, vals, &n);
    if (n_vals) *n_vals = n;
    return JB_OK;
}

extern "C" int jb_quality_tables(int quality, uint32_t ql[64], uint32_t qc[64]) {
    // IJG scaling of the reference's q50 tables (utils.hpp:42-62 = T.81 K.1/K.2)
    static const uint32_t l50[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                     14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                     18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                     49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
    static const uint32_t c50[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                     24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                     99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};
    if (!ql || !qc) return JB_E_INVALID;
    quality = quality < 1 ? 1 : quality > 100 ? 100 : quality;
    int s = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    for (int i = 0; i < 64; ++i) {
        long a = ((long)l50[i] * s + 50) / 100, b = ((long)c50[i] * s + 50) / 100;
        ql[i] = (uint32_t)(a < 1 ? 1 : a > 255 ? 255 : a);
        qc[i] = (uint32_t)(b < 1 ? 1 : b > 255 ? 255 : b);
    }
    return JB_OK;
}
