/* TEST INFRASTRUCTURE ONLY (oracle/) -- see jpeg_oracle.h for scope and pinning.
 *
 * Plain-C restatement of the reference's CPU JPEG path.  Citations are
 * file:line into /root/reference/src.  Build: oracle/Makefile (gcc -O2
 * -ffp-contract=off so that double arithmetic is evaluated exactly as the
 * reference's x86-64 build evaluates it: IEEE binary64, no fused ops).
 */
#include "jpeg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ tables */

/* utils.hpp:42-51 (T.81 Annex K.1), row-major [v][u]. */
const unsigned orc_q50_lum[64] = {
    16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
    14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
    18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
    49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
/* utils.hpp:53-62 (Annex K.2). */
const unsigned orc_q50_chrom[64] = {
    17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
    24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
    99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

/* T.81 Annex K.3 BITS / HUFFVAL lists.  The reference spells the resulting
 * codes out as '0'/'1' strings (huffman.hpp:9,26,43,250); tests check that
 * the codes generated here equal those strings symbol by symbol. */
static const uint8_t bits_dc_lum[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t bits_dc_chr[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t val_dc[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t bits_ac_lum[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t val_ac_lum[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71,
    0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72,
    0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37,
    0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59,
    0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
    0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t bits_ac_chr[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t val_ac_chr[162] = {
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22,
    0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1,
    0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36,
    0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58,
    0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a,
    0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a,
    0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba,
    0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

typedef struct {
    uint32_t code[256];
    uint8_t len[256];
} hufftab;

static hufftab g_tab[4]; /* 0 DC lum, 1 DC chr, 2 AC lum, 3 AC chr */
/* the BITS / HUFFVAL lists in force (Annex K unless orc_set_huffman_specs replaced them), same index */
static uint8_t g_spec_bits[4][16], g_spec_vals[4][256];
static int g_spec_n[4];
uint8_t orc_zigzag_order[64];
static double g_cos[8][8]; /* g_cos[u][x] = cos((2x+1)*u*pi/16) */
static int g_init_done = 0;

/* T.81 Annex C: canonical code assignment from BITS/HUFFVAL. */
static void build_table(hufftab *t, const uint8_t bits[16], const uint8_t *vals) {
    memset(t, 0, sizeof(*t));
    uint32_t code = 0;
    int k = 0;
    for (int l = 1; l <= 16; ++l) {
        for (int i = 0; i < bits[l - 1]; ++i) {
            t->code[vals[k]] = code++;
            t->len[vals[k]] = (uint8_t)l;
            ++k;
        }
        code <<= 1;
    }
}

void orc_init(void) {
    if (g_init_done) return;
    orc_reset_huffman_specs();
    /* zigzag order by the diagonal walk of utils.cpp:539-551 */
    unsigned idx = 0;
    for (int diag = 0; diag < 15; ++diag) {
        int lo = diag - 7 > 0 ? diag - 7 : 0;
        int span = diag < 14 - diag ? diag : 14 - diag;
        for (int i = lo; i <= lo + span; ++i) {
            int row = (diag & 1) ? i : diag - i;
            int col = (diag & 1) ? diag - i : i;
            orc_zigzag_order[idx++] = (uint8_t)(row * 8 + col);
        }
    }
    /* arguments formed exactly as utils.cpp:330 forms them */
    for (size_t u = 0; u < 8; ++u)
        for (size_t x = 0; x < 8; ++x) g_cos[u][x] = cos((2 * x + 1) * u * M_PI / 16.0);
    g_init_done = 1;
}

int orc_table_code(int table, int run, int cat, int typo, uint32_t *code) {
    orc_init();
    int sym = (table < 2) ? cat : ((run << 4) | cat);
    uint32_t c = g_tab[table].code[sym];
    int l = g_tab[table].len[sym];
    /* huffman.hpp:92-98: luma AC 3/4 .. 3/A carry an extra leading '1' */
    if (typo && table == 2 && run == 3 && cat >= 4 && cat <= 10 && l == 16) {
        c |= 1u << 16;
        l = 17;
    }
    if (code) *code = c;
    return l;
}

/* ------------------------------------------------------- per-stage [pinned] */

/* utils.cpp:92-110.  Doubles, left to right, C cast truncates. */
void orc_csc(uint8_t *px, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        uint8_t r = px[3 * i], g = px[3 * i + 1], b = px[3 * i + 2];
        px[3 * i] = (uint8_t)(0.299 * r + 0.587 * g + 0.114 * b);
        px[3 * i + 1] = (uint8_t)(-0.168736 * r - 0.331264 * g + 0.5 * b + 128);
        px[3 * i + 2] = (uint8_t)(0.5 * r - 0.418688 * g - 0.081312 * b + 128);
    }
}

/* utils.cpp:113-141.  2x2 mean of Cb and of Cr written back to all four
 * pixels; an odd last row/column is left untouched. */
void orc_cds(uint8_t *px, size_t W, size_t H) {
    for (size_t y = 0; y + 1 < H; y += 2) {
        for (size_t x = 0; x + 1 < W; x += 2) {
            uint8_t *p[4] = {px + 3 * (y * W + x), px + 3 * (y * W + x + 1), px + 3 * ((y + 1) * W + x),
                             px + 3 * ((y + 1) * W + x + 1)};
            for (int c = 1; c <= 2; ++c) {
                uint8_t m = (uint8_t)((p[0][c] + p[1][c] + p[2][c] + p[3][c]) / 4.0);
                for (int k = 0; k < 4; ++k) p[k][c] = m;
            }
        }
    }
}

/* utils.cpp:184-187 (and cpp:93-98: unchanged when already multiples). */
void orc_padded_size(size_t W, size_t H, size_t mult, size_t *nW, size_t *nH) {
    *nW = (W + mult - 1) / mult * mult;
    *nH = (H + mult - 1) / mult * mult;
}

/* utils.cpp:199-208 then 211-233: copy, mirror right (rows < H), then mirror
 * bottom over the full new width. */
int orc_pad_mirror(const uint8_t *src, size_t W, size_t H, uint8_t *dst, size_t nW, size_t nH) {
    if (nW - W > W || nH - H > H) return -1; /* the reference would index before the image */
    for (size_t y = 0; y < H; ++y) memcpy(dst + 3 * y * nW, src + 3 * y * W, 3 * W);
    for (size_t y = 0; y < H; ++y)
        for (size_t x = W; x < nW; ++x) memcpy(dst + 3 * (y * nW + x), dst + 3 * (y * nW + (W - (x - W + 1))), 3);
    for (size_t y = H; y < nH; ++y)
        for (size_t x = 0; x < nW; ++x) memcpy(dst + 3 * (y * nW + x), dst + 3 * ((H - (y - H + 1)) * nW + x), 3);
    return 0;
}

void orc_u8_to_double(const uint8_t *src, double *dst, size_t n) { /* utils.cpp:236-246 */
    for (size_t i = 0; i < n; ++i) dst[i] = (double)src[i];
}

void orc_subtract(double *img, size_t n, double val) { /* utils.cpp:190-196 */
    for (size_t i = 0; i < n; ++i) img[i] -= val;
}

/* utils.cpp:314-347 for one channel of one block.  blk is [y][x] on input and
 * [v][u] on output.  Output loop u outer / v inner, sum y outer / x inner,
 * term = (sample*cos_x)*cos_y, scale (aU*aV/4) applied last.  inplace != 0
 * reproduces the overwrite-while-reading defect (Q1). */
static void dct_block(double blk[64], int inplace) {
    double src_copy[64];
    const double *in = blk;
    if (!inplace) {
        memcpy(src_copy, blk, sizeof(src_copy));
        in = src_copy;
    }
    for (size_t u = 0; u < 8; ++u) {
        for (size_t v = 0; v < 8; ++v) {
            double au = (u == 0) ? 1.0 / sqrt(2) : 1.0;
            double av = (v == 0) ? 1.0 / sqrt(2) : 1.0;
            double s = 0.0;
            for (size_t y = 0; y < 8; ++y)
                for (size_t x = 0; x < 8; ++x) s += in[y * 8 + x] * g_cos[u][x] * g_cos[v][y];
            s *= (au * av / 4.0);
            blk[v * 8 + u] = s;
        }
    }
}

/* utils.cpp:262-270: every 8x8 block of the AoS double image, 3 channels. */
void orc_dct_image(double *img, size_t W, size_t H, int inplace) {
    orc_init();
    double blk[64];
    for (size_t y0 = 0; y0 < H; y0 += 8)
        for (size_t x0 = 0; x0 < W; x0 += 8)
            for (int c = 0; c < 3; ++c) {
                for (int j = 0; j < 8; ++j)
                    for (int i = 0; i < 8; ++i) blk[j * 8 + i] = img[3 * ((y0 + j) * W + x0 + i) + c];
                dct_block(blk, inplace);
                for (int j = 0; j < 8; ++j)
                    for (int i = 0; i < 8; ++i) img[3 * ((y0 + j) * W + x0 + i) + c] = blk[j * 8 + i];
            }
}

/* utils.cpp:454-467: round half away from zero of F / q[v][u]. */
void orc_quantize_image(double *img, size_t W, size_t H, const unsigned ql[64], const unsigned qc[64]) {
    for (size_t y = 0; y < H; ++y)
        for (size_t x = 0; x < W; ++x) {
            double *p = img + 3 * (y * W + x);
            size_t k = (y % 8) * 8 + (x % 8);
            p[0] = round(p[0] / ql[k]);
            p[1] = round(p[1] / qc[k]);
            p[2] = round(p[2] / qc[k]);
        }
}

/* utils.cpp:482-498: planar by channel, raster block order, row-major in block. */
void orc_blockify(const double *img, size_t W, size_t H, int32_t *linear) {
    size_t rpc = W * H / 64, bx = W / 8;
    for (size_t y = 0; y < H; ++y)
        for (size_t x = 0; x < W; ++x) {
            size_t blk = (y / 8) * bx + x / 8, k = (y % 8) * 8 + x % 8;
            for (int c = 0; c < 3; ++c) linear[(blk + rpc * c) * 64 + k] = (int)img[3 * (y * W + x) + c];
        }
}

void orc_zigzag(const int32_t *linear, int32_t *zz, size_t rows) { /* utils.cpp:539-558 */
    orc_init();
    for (size_t r = 0; r < rows; ++r)
        for (int k = 0; k < 64; ++k) zz[r * 64 + k] = linear[r * 64 + orc_zigzag_order[k]];
}

/* ---- the planar uint32 image of the reference's OpenCL half ------------------------------- */
/* copyImageToVector, utils.cpp:700-707: v[idx] = r, v[idx + n] = g, v[idx + 2n] = b */
void orc_aos_to_planar_u32(const uint8_t *px, size_t W, size_t H, uint32_t *v) {
    size_t n = W * H;
    for (size_t idx = 0; idx < n; ++idx) {
        v[idx] = px[3 * idx];
        v[idx + n] = px[3 * idx + 1];
        v[idx + 2 * n] = px[3 * idx + 2];
    }
}
/* switchVectorChannelOrdering, utils.cpp:745-754: out[3y] = in[y], out[3y+1] = in[y + n], out[3y+2] = in[y + 2n] */
void orc_planar_u32_interleave(const uint32_t *in, size_t W, size_t H, uint32_t *out) {
    size_t n = W * H;
    for (size_t y = 0; y < n; ++y) {
        out[y * 3] = in[y];
        out[y * 3 + 1] = in[y + n];
        out[y * 3 + 2] = in[y + 2 * n];
    }
}

/* utils.cpp:572-609.  pairs needs room for 2*64 ints; returns ints written. */
size_t orc_rle_block(const int32_t zz[64], int32_t *pairs, int always_eob) {
    int last = 0;
    for (int i = 63; i >= 0; --i)
        if (zz[i] != 0) {
            last = i;
            break;
        }
    size_t n = 0;
    int run = 0;
    for (int i = 1; i <= last; ++i) {
        if (zz[i] == 0) {
            if (run == 15) {
                pairs[n++] = 15;
                pairs[n++] = 0;
                run = 0;
            } else {
                ++run;
            }
        } else {
            pairs[n++] = run;
            pairs[n++] = zz[i];
            run = 0;
        }
    }
    if (always_eob || last != 63) { /* Q3: the reference always appends (0,0) */
        pairs[n++] = 0;
        pairs[n++] = 0;
    }
    return n;
}

int orc_category(int v) { /* utils.cpp:623-627: floor(log2|v|)+1 */
    int a = v < 0 ? -v : v, c = 0;
    while (a) {
        ++c;
        a >>= 1;
    }
    return c;
}

int orc_value_bits(int v, uint32_t *bits) { /* utils.cpp:630-653 */
    int c = orc_category(v);
    if (v < 0) v += (1 << c) - 1;
    *bits = (uint32_t)v;
    return c;
}

/* --------------------------------------------------------------- bit sink */

typedef struct {
    uint8_t *buf;
    size_t cap;
    uint64_t nbits;
    int overflow;
} bitsink;

static void put_bits(bitsink *s, uint32_t code, int len) {
    for (int i = len - 1; i >= 0; --i) {
        size_t byte = (size_t)(s->nbits >> 3);
        if (byte >= s->cap) {
            s->overflow = 1;
            ++s->nbits;
            continue;
        }
        if ((s->nbits & 7) == 0) s->buf[byte] = 0;
        if ((code >> i) & 1) s->buf[byte] |= (uint8_t)(0x80 >> (s->nbits & 7));
        ++s->nbits;
    }
}

/* One block as HuffmanEncoder codes it (utils.cpp:667-694): DC category code +
 * value bits, then each RLE pair as AC[run][cat] + value bits. */
static int encode_block(bitsink *s, const int32_t zz[64], int dc_pred, int chroma, int quirks) {
    int typo = (quirks & ORC_Q2_TYPO_TABLES) != 0;
    uint32_t code, vb;
    int diff = zz[0] - dc_pred;
    int cat = orc_value_bits(diff, &vb);
    if (cat > 11) return -1;
    int len = orc_table_code(chroma ? 1 : 0, 0, cat, typo, &code);
    put_bits(s, code, len);
    put_bits(s, vb, cat);
    int32_t pairs[130];
    size_t n = orc_rle_block(zz, pairs, (quirks & ORC_Q3_ALWAYS_EOB) != 0);
    for (size_t j = 0; j < n; j += 2) {
        cat = orc_value_bits(pairs[j + 1], &vb);
        if (cat > 10) return -1; /* the reference tables have 11 columns */
        len = orc_table_code(chroma ? 3 : 2, pairs[j], cat, typo, &code);
        if (len == 0) return -1;
        put_bits(s, code, len);
        put_bits(s, vb, cat);
    }
    return 0;
}

/* utils.cpp:656-698: block i of Y, Cb, Cr interleaved; predictors start at 0
 * and are never reset. */
uint64_t orc_huffman_ref(const int32_t *zz, size_t rpc, int quirks, uint8_t *packed, size_t cap) {
    orc_init();
    bitsink s = {packed, cap, 0, 0};
    int pred[3] = {0, 0, 0};
    for (size_t i = 0; i < rpc; ++i)
        for (int c = 0; c < 3; ++c) {
            const int32_t *b = zz + (i + rpc * c) * 64;
            if (encode_block(&s, b, pred[c], c != 0, quirks)) return (uint64_t)-1;
            pred[c] = b[0];
        }
    return s.nbits;
}

/* ---------------------------------------------------- new surface [unpinned] */

/* IJG quality scaling applied to the reference's q50 tables (utils.hpp:42-62). */
void orc_quality_tables(int quality, unsigned ql[64], unsigned qc[64]) {
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    int s = quality < 50 ? 5000 / quality : 200 - 2 * quality;
    for (int i = 0; i < 64; ++i) {
        long a = ((long)orc_q50_lum[i] * s + 50) / 100, b = ((long)orc_q50_chrom[i] * s + 50) / 100;
        ql[i] = (unsigned)(a < 1 ? 1 : a > 255 ? 255 : a);
        qc[i] = (unsigned)(b < 1 ? 1 : b > 255 ? 255 : b);
    }
}

int orc_blocks_per_mcu(int sub) { return sub == ORC_SUB_420 ? 6 : 3; }

size_t orc_num_mcus(size_t W, size_t H, int sub) {
    size_t m = sub == ORC_SUB_420 ? 16 : 8, nW, nH;
    orc_padded_size(W, H, m, &nW, &nH);
    return (nW / m) * (nH / m);
}

/* Driver order cpp:59-120: CSC, CDS on the unpadded image, then mirror pad. */
int orc_ycc_padded(const uint8_t *rgb, size_t W, size_t H, int sub, uint8_t *dst, size_t *nW, size_t *nH) {
    size_t m = sub == ORC_SUB_420 ? 16 : 8;
    orc_padded_size(W, H, m, nW, nH);
    uint8_t *tmp = (uint8_t *)malloc(W * H * 3);
    if (!tmp) return -1;
    memcpy(tmp, rgb, W * H * 3);
    orc_csc(tmp, W * H);
    if (sub != ORC_SUB_444) orc_cds(tmp, W, H);
    int rc = orc_pad_mirror(tmp, W, H, dst, *nW, *nH);
    free(tmp);
    return rc;
}

/* Level shift, DCT, quantise, zigzag of one 8x8 block of 8-bit samples. */
static void block_to_coef(const uint8_t *smp, size_t stride, size_t step, const unsigned q[64], int inplace,
                          int16_t out[64]) {
    double blk[64];
    for (int j = 0; j < 8; ++j)
        for (int i = 0; i < 8; ++i) {
            blk[j * 8 + i] = (double)smp[(size_t)j * stride + (size_t)i * step]; /* utils.cpp:236 */
            blk[j * 8 + i] -= 128.0;                                           /* utils.cpp:190 */
        }
    dct_block(blk, inplace);                                                   /* utils.cpp:314 */
    for (int k = 0; k < 64; ++k) {
        double r = round(blk[orc_zigzag_order[k]] / q[orc_zigzag_order[k]]);   /* utils.cpp:454 */
        out[k] = (int16_t)(int)r;                                              /* utils.cpp:482 */
    }
}

/* The stages after performCDS for an image that is ALREADY Y,Cb,Cr (AoS bytes, full resolution, unpadded): mirror
 * padding (utils.cpp:199-233), level shift, DCT, quantisation, zigzag -- orc_transform minus orc_csc / orc_cds.  This is
 * what an NV12-style input goes through (its chroma replicated to full resolution is exactly the plane performCDS
 * leaves behind); tests pin it through the identity orc_transform(rgb) == orc_transform_ycc(CSC + CDS of rgb). */
static int transform_padded(uint8_t *ycc, size_t nW, size_t nH, int sub, const unsigned ql[64], const unsigned qc[64],
                            int quirks, int16_t *coef);
int orc_transform_ycc(const uint8_t *ycc_in, size_t W, size_t H, int sub, const unsigned ql[64], const unsigned qc[64],
                      int quirks, int16_t *coef) {
    orc_init();
    size_t nW, nH, m = sub == ORC_SUB_420 ? 16 : 8;
    orc_padded_size(W, H, m, &nW, &nH);
    uint8_t *ycc = (uint8_t *)malloc(nW * nH * 3);
    if (!ycc) return -1;
    if (orc_pad_mirror(ycc_in, W, H, ycc, nW, nH)) {
        free(ycc);
        return -1;
    }
    return transform_padded(ycc, nW, nH, sub, ql, qc, quirks, coef);
}

int orc_transform(const uint8_t *rgb, size_t W, size_t H, int sub, const unsigned ql[64], const unsigned qc[64],
                  int quirks, int16_t *coef) {
    orc_init();
    size_t nW, nH, m = sub == ORC_SUB_420 ? 16 : 8;
    orc_padded_size(W, H, m, &nW, &nH);
    uint8_t *ycc = (uint8_t *)malloc(nW * nH * 3);
    if (!ycc) return -1;
    if (orc_ycc_padded(rgb, W, H, sub, ycc, &nW, &nH)) {
        free(ycc);
        return -1;
    }
    return transform_padded(ycc, nW, nH, sub, ql, qc, quirks, coef);
}

/* blocks of the padded Y,Cb,Cr image in scan order; frees ycc */
static int transform_padded(uint8_t *ycc, size_t nW, size_t nH, int sub, const unsigned ql[64], const unsigned qc[64],
                            int quirks, int16_t *coef) {
    size_t m = sub == ORC_SUB_420 ? 16 : 8;
    int inplace = (quirks & ORC_Q1_INPLACE_DCT) != 0;
    size_t mx = nW / m, my = nH / m;
    for (size_t j = 0; j < my; ++j)
        for (size_t i = 0; i < mx; ++i) {
            int16_t *o = coef + (j * mx + i) * (size_t)orc_blocks_per_mcu(sub) * 64;
            if (sub == ORC_SUB_420) {
                /* Y00 Y01 Y10 Y11, then Cb, Cr sampled at (2i,2j) of the replicated planes */
                for (int b = 0; b < 4; ++b)
                    block_to_coef(ycc + 3 * ((j * 16 + (b >> 1) * 8) * nW + i * 16 + (b & 1) * 8), 3 * nW, 3, ql,
                                  inplace, o + b * 64);
                for (int c = 1; c <= 2; ++c)
                    block_to_coef(ycc + 3 * (j * 16 * nW + i * 16) + c, 6 * nW, 6, qc, inplace, o + (3 + c) * 64);
            } else {
                for (int c = 0; c < 3; ++c)
                    block_to_coef(ycc + 3 * (j * 8 * nW + i * 8) + c, 3 * nW, 3, c ? qc : ql, inplace, o + c * 64);
            }
        }
    free(ycc);
    return 0;
}

static void stuff_interval(const uint8_t *src, uint64_t nbits, uint8_t **dst, size_t *n, size_t cap, int *ovf) {
    size_t nbytes = (size_t)((nbits + 7) >> 3);
    for (size_t i = 0; i < nbytes; ++i) {
        uint8_t b = src[i];
        if (i == nbytes - 1 && (nbits & 7)) b |= (uint8_t)(0xFF >> (nbits & 7)); /* pad with 1s */
        if (*n < cap) (*dst)[*n] = b; else *ovf = 1;
        ++*n;
        if (b == 0xFF) {
            if (*n < cap) (*dst)[*n] = 0; else *ovf = 1;
            ++*n;
        }
    }
}

size_t orc_entropy(const int16_t *coef, size_t n_mcu, int sub, int restart_interval, int quirks, int raw_bits,
                   int rst_phase, int final_rst, uint8_t *out, size_t cap, uint64_t *nbits_out) {
    orc_init();
    int bpm = orc_blocks_per_mcu(sub);
    size_t ri = restart_interval > 0 ? (size_t)restart_interval : n_mcu;
    if (raw_bits && restart_interval > 0) return (size_t)-1;
    size_t tmp_cap = ri * (size_t)bpm * 256 + 64; /* > worst case 64*(17+11) bits per block */
    uint8_t *tmp = raw_bits ? out : (uint8_t *)malloc(tmp_cap);
    if (!tmp) return (size_t)-1;
    size_t n = 0, interval = 0;
    int ovf = 0;
    uint64_t total_bits = 0;
    for (size_t m0 = 0; m0 < n_mcu; m0 += ri, ++interval) {
        size_t m1 = m0 + ri < n_mcu ? m0 + ri : n_mcu;
        bitsink s = {tmp, raw_bits ? cap : tmp_cap, 0, 0};
        int pred[3] = {0, 0, 0};
        for (size_t m = m0; m < m1; ++m)
            for (int b = 0; b < bpm; ++b) {
                int comp = sub == ORC_SUB_420 ? (b < 4 ? 0 : b - 3) : b;
                int32_t zz[64];
                const int16_t *cb = coef + (m * (size_t)bpm + (size_t)b) * 64;
                for (int k = 0; k < 64; ++k) zz[k] = cb[k];
                if (encode_block(&s, zz, pred[comp], comp != 0, quirks)) {
                    if (!raw_bits) free(tmp);
                    return (size_t)-1;
                }
                pred[comp] = zz[0];
            }
        total_bits += s.nbits;
        if (s.overflow) ovf = 1;
        if (raw_bits) {
            n = (size_t)((s.nbits + 7) >> 3);
        } else {
            stuff_interval(tmp, s.nbits, &out, &n, cap, &ovf);
            if (m1 < n_mcu || final_rst) { /* RSTm, m = interval index mod 8 (T.81 E.1.4) */
                uint8_t mk[2] = {0xFF, (uint8_t)(0xD0 + ((interval + (size_t)rst_phase) & 7))};
                for (int k = 0; k < 2; ++k) {
                    if (n < cap) out[n] = mk[k]; else ovf = 1;
                    ++n;
                }
            }
        }
    }
    if (!raw_bits) free(tmp);
    if (nbits_out) *nbits_out = total_bits;
    return ovf ? (size_t)-1 : n;
}

static void put_u16(uint8_t *p, unsigned v) {
    p[0] = (uint8_t)(v >> 8);
    p[1] = (uint8_t)v;
}

/* Segment order validated with PIL and OpenCV. */
size_t orc_jfif_header(size_t W, size_t H, int sub, const unsigned ql[64], const unsigned qc[64],
                       int restart_interval, uint8_t *out, size_t cap) {
    orc_init();
    uint8_t h[2048];
    size_t n = 0;
    static const uint8_t app0[] = {0xFF, 0xD8, 0xFF, 0xE0, 0x00, 0x10, 'J', 'F', 'I', 'F', 0x00,
                                   0x01, 0x01, 0x00, 0x00, 0x01, 0x00, 0x01, 0x00, 0x00};
    memcpy(h + n, app0, sizeof(app0));
    n += sizeof(app0);
    for (int t = 0; t < 2; ++t) {
        h[n++] = 0xFF; h[n++] = 0xDB; h[n++] = 0x00; h[n++] = 0x43; h[n++] = (uint8_t)t;
        for (int k = 0; k < 64; ++k) h[n++] = (uint8_t)(t ? qc : ql)[orc_zigzag_order[k]];
    }
    /* SOF0 sizes are 16 bit: 65536 is declared as 65535. */
    unsigned dw = W > 65535 ? 65535u : (unsigned)W, dh = H > 65535 ? 65535u : (unsigned)H;
    h[n++] = 0xFF; h[n++] = 0xC0; h[n++] = 0x00; h[n++] = 0x11; h[n++] = 0x08;
    put_u16(h + n, dh); n += 2;
    put_u16(h + n, dw); n += 2;
    h[n++] = 3;
    h[n++] = 1; h[n++] = (uint8_t)(sub == ORC_SUB_420 ? 0x22 : 0x11); h[n++] = 0;
    h[n++] = 2; h[n++] = 0x11; h[n++] = 1;
    h[n++] = 3; h[n++] = 0x11; h[n++] = 1;
    static const int order[4] = {0, 2, 1, 3}; /* DHT segments: DC lum, AC lum, DC chr, AC chr */
    static const uint8_t tcth[4] = {0x00, 0x10, 0x01, 0x11};
    for (int t = 0; t < 4; ++t) {
        int i = order[t], nv = g_spec_n[i];
        h[n++] = 0xFF; h[n++] = 0xC4;
        put_u16(h + n, (unsigned)(2 + 1 + 16 + nv)); n += 2;
        h[n++] = tcth[t];
        memcpy(h + n, g_spec_bits[i], 16); n += 16;
        memcpy(h + n, g_spec_vals[i], (size_t)nv); n += (size_t)nv;
    }
    if (restart_interval > 0) {
        h[n++] = 0xFF; h[n++] = 0xDD; h[n++] = 0x00; h[n++] = 0x04;
        put_u16(h + n, (unsigned)restart_interval); n += 2;
    }
    static const uint8_t sos[] = {0xFF, 0xDA, 0x00, 0x0C, 0x03, 0x01, 0x00, 0x02, 0x11, 0x03, 0x11, 0x00, 0x3F, 0x00};
    memcpy(h + n, sos, sizeof(sos));
    n += sizeof(sos);
    if (out) memcpy(out, h, n < cap ? n : cap);
    return n;
}

size_t orc_encode_jfif(const uint8_t *rgb, size_t W, size_t H, int sub, const unsigned ql[64],
                       const unsigned qc[64], int restart_interval, int quirks, uint8_t *out, size_t cap) {
    if (restart_interval < 0 || restart_interval > 65535) return (size_t)-1;
    size_t n_mcu = orc_num_mcus(W, H, sub);
    int16_t *coef = (int16_t *)malloc(n_mcu * (size_t)orc_blocks_per_mcu(sub) * 64 * sizeof(int16_t));
    if (!coef) return (size_t)-1;
    if (orc_transform(rgb, W, H, sub, ql, qc, quirks, coef)) {
        free(coef);
        return (size_t)-1;
    }
    size_t n = orc_jfif_header(W, H, sub, ql, qc, restart_interval, out, cap);
    if (n > cap) {
        free(coef);
        return (size_t)-1;
    }
    size_t e = orc_entropy(coef, n_mcu, sub, restart_interval, quirks, 0, 0, 0, out + n, cap - n, NULL);
    free(coef);
    if (e == (size_t)-1 || n + e + 2 > cap) return (size_t)-1;
    n += e;
    out[n++] = 0xFF;
    out[n++] = 0xD9;
    return n;
}

/* ---- optimised Huffman tables [unpinned by the reference: it has fixed tables] ----
 * Two passes like libjpeg's optimize_coding: count the symbols the coder would emit, derive BITS / HUFFVAL
 * with the procedure of T.81 Annex K.2 (Figures K.1-K.4) in the form libjpeg's jpeg_gen_optimal_table gives it
 * (a reserved 257th symbol keeps the all-ones code unused, ties go to the larger symbol, lengths limited to 16),
 * then code with those tables and write them into the DHT segments. */
void orc_reset_huffman_specs(void) {
    static const uint8_t *bitsv[4] = {bits_dc_lum, bits_dc_chr, bits_ac_lum, bits_ac_chr};
    static const uint8_t *valsv[4] = {val_dc, val_dc, val_ac_lum, val_ac_chr};
    for (int i = 0; i < 4; ++i) {
        g_spec_n[i] = i < 2 ? 12 : 162;
        memcpy(g_spec_bits[i], bitsv[i], 16);
        memset(g_spec_vals[i], 0, 256);
        memcpy(g_spec_vals[i], valsv[i], (size_t)g_spec_n[i]);
        build_table(&g_tab[i], g_spec_bits[i], g_spec_vals[i]);
    }
}

/* index: 0 DC lum, 1 DC chr, 2 AC lum, 3 AC chr */
void orc_set_huffman_specs(const uint8_t bits[4][16], const uint8_t vals[4][256]) {
    orc_init();
    for (int i = 0; i < 4; ++i) {
        int n = 0;
        for (int l = 0; l < 16; ++l) n += bits[i][l];
        g_spec_n[i] = n;
        memcpy(g_spec_bits[i], bits[i], 16);
        memcpy(g_spec_vals[i], vals[i], 256);
        build_table(&g_tab[i], g_spec_bits[i], g_spec_vals[i]);
    }
}

/* Symbols the entropy coder emits for these coefficients (same walk as orc_entropy). */
void orc_symbol_histogram(const int16_t *coef, size_t n_mcu, int sub, int restart_interval, int quirks,
                          uint64_t hist[4][256]) {
    int bpm = orc_blocks_per_mcu(sub);
    size_t ri = restart_interval > 0 ? (size_t)restart_interval : n_mcu;
    memset(hist, 0, 4 * 256 * sizeof(uint64_t));
    for (size_t m0 = 0; m0 < n_mcu; m0 += ri) {
        size_t m1 = m0 + ri < n_mcu ? m0 + ri : n_mcu;
        int pred[3] = {0, 0, 0};
        for (size_t m = m0; m < m1; ++m)
            for (int b = 0; b < bpm; ++b) {
                int comp = sub == ORC_SUB_420 ? (b < 4 ? 0 : b - 3) : b, chroma = comp != 0;
                int32_t zz[64], pairs[130];
                const int16_t *cb = coef + (m * (size_t)bpm + (size_t)b) * 64;
                for (int k = 0; k < 64; ++k) zz[k] = cb[k];
                hist[chroma][orc_category(zz[0] - pred[comp])]++;
                pred[comp] = zz[0];
                size_t n = orc_rle_block(zz, pairs, (quirks & ORC_Q3_ALWAYS_EOB) != 0);
                for (size_t j = 0; j < n; j += 2) hist[2 + chroma][(pairs[j] << 4) | orc_category(pairs[j + 1])]++;
            }
    }
}

/* T.81 K.2 / libjpeg jpeg_gen_optimal_table.  Returns the number of symbols (HUFFVAL entries). */
int orc_optimal_spec(const uint64_t freq_in[256], uint8_t bits_out[16], uint8_t vals_out[256]) {
    enum { MAXLEN = 32 };
    uint64_t base[256], freq[257];
    int codesize[257], others[257], bits[MAXLEN + 1];
    for (int i = 0; i < 256; ++i) base[i] = freq_in[i];
    for (;;) {
        for (int i = 0; i < 256; ++i) freq[i] = base[i];
        freq[256] = 1; /* reserved: guarantees that no real symbol gets the all-ones code */
        for (int i = 0; i < 257; ++i) { codesize[i] = 0; others[i] = -1; }
        memset(bits, 0, sizeof(bits));
        for (;;) {
            int c1 = -1, c2 = -1;
            uint64_t v = UINT64_MAX;
            for (int i = 0; i <= 256; ++i)
                if (freq[i] && freq[i] <= v) { v = freq[i]; c1 = i; } /* smallest, larger symbol on ties */
            v = UINT64_MAX;
            for (int i = 0; i <= 256; ++i)
                if (freq[i] && freq[i] <= v && i != c1) { v = freq[i]; c2 = i; }
            if (c2 < 0) break;
            freq[c1] += freq[c2];
            freq[c2] = 0;
            codesize[c1]++;
            while (others[c1] >= 0) { c1 = others[c1]; codesize[c1]++; }
            others[c1] = c2;
            codesize[c2]++;
            while (others[c2] >= 0) { c2 = others[c2]; codesize[c2]++; }
        }
        int too_deep = 0;
        for (int i = 0; i <= 256; ++i)
            if (codesize[i] > MAXLEN) too_deep = 1;
        if (!too_deep) break;
        /* libjpeg gives up here (JERR_HUFF_CLEN_OVERFLOW; needs > 9 million symbols in a Fibonacci-like
         * distribution).  Flatten the distribution instead: halve every count, rounding up, and start over. */
        for (int i = 0; i < 256; ++i)
            if (base[i]) base[i] = (base[i] + 1) / 2;
    }
    for (int i = 0; i <= 256; ++i)
        if (codesize[i]) bits[codesize[i]]++;
    int i;
    for (i = MAXLEN; i > 16; --i) /* Figure K.3: move the deepest pairs up until no code is longer than 16 */
        while (bits[i] > 0) {
            int j = i - 2;
            while (bits[j] == 0) --j;
            bits[i] -= 2;
            bits[i - 1]++;
            bits[j + 1] += 2;
            bits[j]--;
        }
    while (bits[i] == 0) --i;
    bits[i]--; /* the reserved symbol had the longest code */
    int n = 0;
    for (int l = 1; l <= 16; ++l) bits_out[l - 1] = (uint8_t)bits[l];
    memset(vals_out, 0, 256);
    for (int l = 1; l <= MAXLEN; ++l) /* Figure K.4: symbols sorted by code size, then by value */
        for (int j = 0; j < 256; ++j)
            if (codesize[j] == l) vals_out[n++] = (uint8_t)j;
    return n;
}

size_t orc_encode_jfif_optimized(const uint8_t *rgb, size_t W, size_t H, int sub, const unsigned ql[64],
                                 const unsigned qc[64], int restart_interval, int quirks, uint8_t *out, size_t cap) {
    if (restart_interval < 0 || restart_interval > 65535) return (size_t)-1;
    orc_init();
    size_t n_mcu = orc_num_mcus(W, H, sub);
    int16_t *coef = (int16_t *)malloc(n_mcu * (size_t)orc_blocks_per_mcu(sub) * 64 * sizeof(int16_t));
    if (!coef) return (size_t)-1;
    if (orc_transform(rgb, W, H, sub, ql, qc, quirks, coef)) {
        free(coef);
        return (size_t)-1;
    }
    uint64_t hist[4][256];
    uint8_t bits[4][16], vals[4][256];
    orc_symbol_histogram(coef, n_mcu, sub, restart_interval, quirks, hist);
    for (int i = 0; i < 4; ++i)
        if (orc_optimal_spec(hist[i], bits[i], vals[i]) < 0) {
            free(coef);
            return (size_t)-1;
        }
    orc_set_huffman_specs(bits, vals);
    size_t n = orc_jfif_header(W, H, sub, ql, qc, restart_interval, out, cap), e = (size_t)-1;
    if (n <= cap) e = orc_entropy(coef, n_mcu, sub, restart_interval, quirks & ~ORC_Q2_TYPO_TABLES, 0, 0, 0, out + n, cap - n, NULL);
    orc_reset_huffman_specs();
    free(coef);
    if (e == (size_t)-1 || n + e + 2 > cap) return (size_t)-1;
    n += e;
    out[n++] = 0xFF;
    out[n++] = 0xD9;
    return n;
}

/* SURVEY.md section 8d generator: per channel a triangle wave of period
 * 97/61/41 px (amplitude 96 around 128) plus uniform +-4 noise from
 * splitmix64; integer-only so that CPU and GPU agree bit for bit. */
static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

void orc_synth_rgb(uint64_t seed, size_t W, size_t y0, size_t rows, uint8_t *out) {
    static const unsigned P[3] = {97, 61, 41}, A[3] = {1, 1, 2}, B[3] = {1, 2, 1};
    for (size_t y = y0; y < y0 + rows; ++y)
        for (size_t x = 0; x < W; ++x)
            for (unsigned c = 0; c < 3; ++c) {
                unsigned ph = (unsigned)((x * A[c] + y * B[c]) % P[c]);
                unsigned v = ph <= P[c] / 2 ? ph : P[c] - ph;
                int base = 32 + (int)(v * 192 / (P[c] / 2));
                uint64_t h = splitmix64(seed ^ ((((uint64_t)y << 32) | (uint64_t)x) * 3 + c));
                int val = base + (int)((h >> 56) % 9) - 4;
                out[((y - y0) * W + x) * 3 + c] = (uint8_t)(val < 0 ? 0 : val > 255 ? 255 : val);
            }
}
