// Decode path: baseline JFIF -> coefficients -> RGB8 on the GPU, so that the PSNR loop of the
// parity report (north_star: "decoded PSNR of both encoders must agree within 0.01 dB") needs no CPU decoder at
// gigapixel sizes, and so that the byte stream can be checked by a round trip (decoded coefficients == the
// coefficients the encoder coded).  The reference has no decoder; the reconstruction follows the decoder every
// independent check in this repository uses -- libjpeg(-turbo), as PIL and OpenCV link it -- operation for operation, so
// that decoded pixels are IDENTICAL to theirs (tests/test_gpu_decode.py):
//   entropy decode   T.81 F.2.2 (Huffman, run/size symbols, EOB/ZRL, DC prediction, FF00 unstuffing, RSTn)
//   dequantisation   coefficient x table entry
//   IDCT             the "slow-but-accurate" integer IDCT (jidctint.c: 13-bit constants, 2 extra bits after pass 1)
//   upsampling       h2v2 "fancy" (triangle filter, jdsample.c) for 4:2:0
//   colour           integer YCbCr -> RGB with 16-bit fixed-point tables (jdcolor.c)
// Parallelism: one thread per restart interval for the entropy decode (an interval is the unit that can be decoded
// independently; a file without DRI is one interval), one thread per 8x8 block for the IDCT, one per pixel after that.
#include <string.h>

#include "jb_pixels.cuh"

namespace jb {

// ---------------------------------------------------------------- host: marker parser --
static uint32_t be16(const uint8_t* p) { return ((uint32_t)p[0] << 8) | p[1]; }

// Huffman decoding tables of one DHT table: 9-bit look-ahead (length << 8 | symbol) + the canonical-code arrays
static void build_dec_table(const uint8_t bits[16], const uint8_t* vals, DecTable* t) {
    memset(t, 0, sizeof(*t));
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        t->valptr[l] = (uint16_t)k;
        t->mincode[l] = code;
        for (int i = 0; i < bits[l - 1]; ++i, ++k, ++code) {
            t->vals[k] = vals[k];
            if (l <= 9)
                for (int f = 0; f < (1 << (9 - l)); ++f) t->look[(code << (9 - l)) | f] = (uint16_t)((l << 8) | vals[k]);
        }
        t->maxcode[l] = bits[l - 1] ? code - 1 : -1;
        code <<= 1;
    }
    t->maxcode[17] = 0x7FFFFFFF;
}

// Baseline, 8 bit, three components, luma sampled 1x1 (4:4:4) or 2x2 (4:2:0) with 1x1 chroma, one interleaved scan.
int parse_jfif(const uint8_t* d, size_t n, JfifInfo* o, DecTables* tabs) {
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return JB_E_INVALID;
    memset(o, 0, sizeof(*o));
    uint8_t bits[4][16], vals[4][256];
    bool have_tab[4] = {false, false, false, false}, have_q[4] = {false, false, false, false}, have_sof = false;
    uint32_t q[4][64];
    int comp_q[3] = {0, 1, 1}, comp_id[3] = {1, 2, 3};
    size_t p = 2;
    while (p + 4 <= n) {
        if (d[p] != 0xFF) return JB_E_INVALID;
        const uint8_t m = d[p + 1];
        if (m == 0xFF) { ++p; continue; }  // fill byte
        const size_t len = be16(d + p + 2);
        if (len < 2 || p + 2 + len > n) return JB_E_NOSPACE;  // the caller has not handed over the whole header yet
        const uint8_t* s = d + p + 4;
        const size_t sl = len - 2;
        if (m == 0xDB) {  // DQT (zigzag order in the file)
            size_t i = 0;
            while (i < sl) {
                const int pq = s[i] >> 4, tq = s[i] & 15;
                if (pq != 0 || tq > 3 || i + 65 > sl) return JB_E_UNSUPPORTED;
                for (int k = 0; k < 64; ++k) q[tq][kZigzag[k]] = s[i + 1 + k];
                have_q[tq] = true;
                i += 65;
            }
        } else if (m == 0xC0) {  // SOF0
            if (sl < 15 || s[0] != 8 || s[5] != 3) return JB_E_UNSUPPORTED;
            o->H = be16(s + 1);
            o->W = be16(s + 3);
            for (int c = 0; c < 3; ++c) {
                comp_id[c] = s[6 + 3 * c];
                const int hv = s[7 + 3 * c];
                comp_q[c] = s[8 + 3 * c];
                if (c == 0) {
                    if (hv == 0x11) o->sub = JB_SUB_444;
                    else if (hv == 0x22) o->sub = JB_SUB_420;
                    else return JB_E_UNSUPPORTED;
                } else if (hv != 0x11) {
                    return JB_E_UNSUPPORTED;
                }
            }
            have_sof = true;
        } else if (m == 0xC4) {  // DHT
            size_t i = 0;
            while (i + 17 <= sl) {
                const int tc = s[i] >> 4, th = s[i] & 15;
                if (tc > 1 || th > 1) return JB_E_UNSUPPORTED;
                int cnt = 0;
                for (int k = 0; k < 16; ++k) cnt += s[i + 1 + k];
                if (cnt > 256 || i + 17 + cnt > sl) return JB_E_INVALID;
                memcpy(bits[tc * 2 + th], s + i + 1, 16);
                memcpy(vals[tc * 2 + th], s + i + 17, cnt);
                have_tab[tc * 2 + th] = true;
                i += 17 + cnt;
            }
        } else if (m == 0xDD) {  // DRI
            if (sl < 2) return JB_E_INVALID;
            o->restart_interval = be16(s);
        } else if (m == 0xDA) {  // SOS: the entropy-coded data follows
            if (!have_sof || sl < 10 || s[0] != 3) return JB_E_UNSUPPORTED;
            for (int c = 0; c < 3; ++c) {
                if (s[1 + 2 * c] != comp_id[c]) return JB_E_UNSUPPORTED;
                o->dc_tab[c] = s[2 + 2 * c] >> 4;
                o->ac_tab[c] = s[2 + 2 * c] & 15;
                if (o->dc_tab[c] > 1 || o->ac_tab[c] > 1 || !have_tab[o->dc_tab[c]] || !have_tab[2 + o->ac_tab[c]] || comp_q[c] > 3 ||
                    !have_q[comp_q[c]])
                    return JB_E_INVALID;
                memcpy(o->q[c], q[comp_q[c]], sizeof(o->q[c]));
            }
            if (s[7] != 0 || s[8] != 63 || s[9] != 0) return JB_E_UNSUPPORTED;  // not a baseline scan
            o->scan_offset = p + 2 + len;
            for (int t = 0; t < 4; ++t)
                if (have_tab[t]) build_dec_table(bits[t], vals[t], &tabs->t[t]);
            if (o->W == 0 || o->H == 0) return JB_E_UNSUPPORTED;  // (DNL-defined heights are not supported)
            return JB_OK;
        } else if ((m >= 0xC1 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC)) {
            return JB_E_UNSUPPORTED;  // not baseline sequential Huffman
        }
        p += 2 + len;
    }
    return JB_E_NOSPACE;
}

// ---------------------------------------------------------------- restart-interval boundaries --
// RSTn markers split the entropy-coded data into independently decodable intervals.  Two passes: markers per
// 256-byte chunk, then (after a scan) their positions in order.  (0xFF is always followed by 0x00 inside the data.)
__device__ __forceinline__ bool is_rst(const uint8_t* d, size_t i, size_t n) { return i + 1 < n && d[i] == 0xFF && (d[i + 1] & 0xF8) == 0xD0; }

__global__ void k_rst_count(const uint8_t* __restrict__ d, size_t n, uint32_t* __restrict__ cnt) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x, lo = c * 256;
    if (lo >= n) return;
    uint32_t k = 0;
    for (size_t i = lo; i < lo + 256 && i < n; ++i) k += is_rst(d, i, n);
    cnt[c] = k;
}
// exclusive prefix of the chunk counts in two levels: every thread scans a group of 1024 chunks in place, one thread
// scans the group totals (a gigapixel scan has ~3 M chunks, ~3 k groups)
__global__ void k_rst_scan_groups(uint32_t* cnt, size_t n_chunks, uint32_t* group_total) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x, lo = g * 1024;
    if (lo >= n_chunks) return;
    uint32_t run = 0;
    for (size_t i = lo; i < lo + 1024 && i < n_chunks; ++i) {
        const uint32_t k = cnt[i];
        cnt[i] = run;
        run += k;
    }
    group_total[g] = run;
}
__global__ void k_rst_scan_totals(uint32_t* group_total, size_t n_groups) {
    uint32_t run = 0;
    for (size_t i = 0; i < n_groups; ++i) {
        const uint32_t k = group_total[i];
        group_total[i] = run;
        run += k;
    }
}
__global__ void k_rst_positions(const uint8_t* __restrict__ d, size_t n, const uint32_t* __restrict__ base, const uint32_t* __restrict__ group_base,
                                uint64_t* __restrict__ start, uint32_t max_int) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x, lo = c * 256;
    if (lo >= n) return;
    uint32_t k = base[c] + group_base[c >> 10];
    for (size_t i = lo; i < lo + 256 && i < n; ++i)
        if (is_rst(d, i, n)) {
            if (k + 1 < max_int) start[k + 1] = i + 2;  // interval k+1 begins after marker k
            ++k;
        }
}

// ---------------------------------------------------------------- entropy decode --
struct BitReader {
    const uint8_t* p;
    const uint8_t* end;
    uint64_t buf;
    int n;  // valid bits in buf (left aligned at bit 63)
    __device__ __forceinline__ void init(const uint8_t* b, const uint8_t* e) {
        p = b;
        end = e;
        buf = 0;
        n = 0;
    }
    __device__ __forceinline__ void fill() {
        while (n <= 56) {
            uint32_t byte = 0;
            if (p < end) {
                byte = *p;
                if (byte == 0xFF) {
                    if (p + 1 < end && p[1] == 0x00) p += 2;  // stuffed zero
                    else { byte = 0; p = end; }               // a marker ends the interval: feed zeros
                } else {
                    ++p;
                }
            }
            buf |= (uint64_t)byte << (56 - n);
            n += 8;
        }
    }
    __device__ __forceinline__ uint32_t peek(int k) const { return (uint32_t)(buf >> (64 - k)); }
    __device__ __forceinline__ void skip(int k) {
        buf <<= k;
        n -= k;
    }
};

__device__ __forceinline__ int huff_symbol(BitReader& br, const DecTable& t) {
    br.fill();
    const uint32_t e = t.look[br.peek(9)];
    if (e) {
        br.skip((int)(e >> 8));
        return (int)(e & 0xFF);
    }
    int l = 10;
    int code = (int)br.peek(10);
    while (l <= 16 && code > t.maxcode[l]) {
        ++l;
        code = (int)br.peek(l);
    }
    if (l > 16) return 0;  // corrupt data
    br.skip(l);
    return t.vals[t.valptr[l] + code - t.mincode[l]];
}
__device__ __forceinline__ int receive_extend(BitReader& br, int s) {  // T.81 F.2.2.1 RECEIVE + EXTEND
    if (!s) return 0;
    br.fill();
    const int v = (int)br.peek(s);
    br.skip(s);
    return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v;
}

// One thread per restart interval: coefficients in zigzag order, scan layout [mcu][block][64] (int16, pre-zeroed).
__global__ void __launch_bounds__(64) k_huff_decode(const uint8_t* __restrict__ data, size_t n_bytes, const uint64_t* __restrict__ start,
                                                    uint32_t n_int, uint32_t ri, uint32_t n_mcu, int bpm, const DecTables* __restrict__ tabs,
                                                    const __grid_constant__ JfifInfo info, int16_t* __restrict__ coef) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_int) return;
    const uint64_t b0 = start[k], b1 = k + 1 < n_int ? start[k + 1] : n_bytes;
    if (b0 > n_bytes || b1 > n_bytes || b0 > b1) return;
    BitReader br;
    br.init(data + b0, data + b1);
    int pred[3] = {0, 0, 0};
    const uint32_t m0 = k * ri, m1 = min(m0 + ri, n_mcu);
    for (uint32_t m = m0; m < m1; ++m)
        for (int j = 0; j < bpm; ++j) {
            const int c = bpm == 3 ? j : (j < 4 ? 0 : j - 3);
            int16_t* blk = coef + ((size_t)m * bpm + j) * 64;
            const int s = huff_symbol(br, tabs->t[info.dc_tab[c]]);
            pred[c] += receive_extend(br, s);
            blk[0] = (int16_t)pred[c];
            const DecTable& ac = tabs->t[2 + info.ac_tab[c]];
            for (int z = 1; z < 64;) {
                const int rs = huff_symbol(br, ac), r = rs >> 4, sz = rs & 15;
                if (sz == 0) {
                    if (r != 15) break;  // EOB
                    z += 16;             // ZRL
                    continue;
                }
                z += r;
                if (z > 63) break;
                blk[z++] = (int16_t)receive_extend(br, sz);
            }
        }
}

// ---------------------------------------------------------------- IDCT --
#define JD_FIX_0_298631336 2446
#define JD_FIX_0_390180644 3196
#define JD_FIX_0_541196100 4433
#define JD_FIX_0_765366865 6270
#define JD_FIX_0_899976223 7373
#define JD_FIX_1_175875602 9633
#define JD_FIX_1_501321110 12299
#define JD_FIX_1_847759065 15137
#define JD_FIX_1_961570560 16069
#define JD_FIX_2_053119869 16819
#define JD_FIX_2_562915447 20995
#define JD_FIX_3_072711026 25172

typedef long long jlong;  // libjpeg's JLONG is a 64-bit long on LP64 hosts
__device__ __forceinline__ int descale(jlong x, int n) { return (int)((x + ((jlong)1 << (n - 1))) >> n); }
// the IDCT's range-limit table (jdmaster.c prepare_range_limit_table), index = value & 1023, centred on 128
__device__ __forceinline__ uint8_t range_limit_idct(int x) {
    const int v = x & 1023;
    return (uint8_t)(v < 128 ? v + 128 : v < 512 ? 255 : v < 896 ? 0 : v - 896);
}
// one 1-D pass of jidctint.c on in[0..7] (stride 1 here); `shift` = CONST_BITS - PASS1_BITS (pass 1) or CONST_BITS + PASS1_BITS + 3
__device__ __forceinline__ void idct8(const int* in, int* out, int shift) {
    jlong z2 = in[2], z3 = in[6];
    jlong z1 = (z2 + z3) * JD_FIX_0_541196100;
    jlong tmp2 = z1 + z3 * (-JD_FIX_1_847759065), tmp3 = z1 + z2 * JD_FIX_0_765366865;
    z2 = in[0];
    z3 = in[4];
    jlong tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;  // << CONST_BITS
    const jlong tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7];
    tmp1 = in[5];
    tmp2 = in[3];
    tmp3 = in[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    jlong z4 = tmp1 + tmp3;
    const jlong z5 = (z3 + z4) * JD_FIX_1_175875602;
    tmp0 *= JD_FIX_0_298631336;
    tmp1 *= JD_FIX_2_053119869;
    tmp2 *= JD_FIX_3_072711026;
    tmp3 *= JD_FIX_1_501321110;
    z1 *= -JD_FIX_0_899976223;
    z2 *= -JD_FIX_2_562915447;
    z3 *= -JD_FIX_1_961570560;
    z4 *= -JD_FIX_0_390180644;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    out[0] = descale(tmp10 + tmp3, shift);
    out[7] = descale(tmp10 - tmp3, shift);
    out[1] = descale(tmp11 + tmp2, shift);
    out[6] = descale(tmp11 - tmp2, shift);
    out[2] = descale(tmp12 + tmp1, shift);
    out[5] = descale(tmp12 - tmp1, shift);
    out[3] = descale(tmp13 + tmp0, shift);
    out[4] = descale(tmp13 - tmp0, shift);
}

// One thread per block: dequantise, IDCT (columns then rows, as jpeg_idct_islow), store bytes into the component's
// plane: Y at full padded resolution, Cb / Cr at the component's own padded resolution.
__global__ void __launch_bounds__(128) k_idct(const int16_t* __restrict__ coef, uint32_t n_mcu, int mcux, int bpm, const __grid_constant__ JfifInfo info,
                                              uint8_t* __restrict__ py, uint8_t* __restrict__ pcb, uint8_t* __restrict__ pcr, size_t pitch_y,
                                              size_t pitch_c) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n_mcu * bpm) return;
    const uint32_t m = (uint32_t)(t / bpm);
    const int j = (int)(t - (size_t)m * bpm), my = (int)(m / mcux), mx = (int)(m - (uint32_t)my * mcux);
    const int c = bpm == 3 ? j : (j < 4 ? 0 : j - 3);
    uint8_t* dst;
    size_t pitch;
    if (c == 0) {
        pitch = pitch_y;
        dst = bpm == 3 ? py + (size_t)my * 8 * pitch + mx * 8 : py + (size_t)(my * 16 + (j >> 1) * 8) * pitch + mx * 16 + (j & 1) * 8;
    } else {
        pitch = pitch_c;
        dst = (c == 1 ? pcb : pcr) + (size_t)my * 8 * pitch + mx * 8;
    }
    const int16_t* z = coef + t * 64;
    int in[64], ws[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) {
        const int nat = zz_nat(k);
        in[nat] = (int)z[k] * (int)info.q[c][nat];  // DEQUANTIZE
    }
#pragma unroll
    for (int col = 0; col < 8; ++col) {  // pass 1: columns
        int a[8], o[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) a[r] = in[r * 8 + col];
        idct8(a, o, 13 - 2);
#pragma unroll
        for (int r = 0; r < 8; ++r) ws[r * 8 + col] = o[r];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {  // pass 2: rows, then the range limit around 128
        int o[8];
        idct8(ws + r * 8, o, 13 + 2 + 3);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            lo |= (uint32_t)range_limit_idct(o[x]) << (8 * x);
            hi |= (uint32_t)range_limit_idct(o[4 + x]) << (8 * x);
        }
        *reinterpret_cast<uint2*>(dst + (size_t)r * pitch) = make_uint2(lo, hi);
    }
}

// ---------------------------------------------------------------- upsampling + colour --
// jdcolor.c build_ycc_rgb_table, SCALEBITS = 16
__device__ __forceinline__ uint8_t clamp_u8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
__device__ __forceinline__ void ycc_to_rgb(int y, int cb, int cr, uint8_t* o) {
    const int xb = cb - 128, xr = cr - 128;
    o[0] = clamp_u8(y + ((91881 * xr + 32768) >> 16));                         // FIX(1.40200)
    o[1] = clamp_u8(y + ((-22554 * xb + 32768 + (-46802) * xr) >> 16));       // FIX(0.34414), FIX(0.71414)
    o[2] = clamp_u8(y + ((116130 * xb + 32768) >> 16));                        // FIX(1.77200)
}

// h2v2 fancy upsampling (jdsample.c h2v2_fancy_upsample) of one chroma plane at output pixel (x, y): vertical
// 3:1 blend of the nearer and the farther row (edge rows replicated), then the horizontal 3:1 blend with the
// rounding constants 8 (even output columns) / 7 (odd), and the special first / last columns.
__device__ __forceinline__ int fancy_h2v2(const uint8_t* __restrict__ p, size_t pitch, int cw, int ch, int x, int y) {
    const int cy = y >> 1, cx = x >> 1;
    int oy = (y & 1) ? cy + 1 : cy - 1;
    oy = oy < 0 ? 0 : oy >= ch ? ch - 1 : oy;
    const uint8_t *r0 = p + (size_t)cy * pitch, *r1 = p + (size_t)oy * pitch;
    const int cur = 3 * r0[cx] + r1[cx];
    if (x & 1) {
        if (cx == cw - 1) return (cur * 4 + 7) >> 4;
        return (cur * 3 + 3 * r0[cx + 1] + r1[cx + 1] + 7) >> 4;
    }
    if (cx == 0) return (cur * 4 + 8) >> 4;
    return (cur * 3 + 3 * r0[cx - 1] + r1[cx - 1] + 8) >> 4;
}

__global__ void k_color(const uint8_t* __restrict__ py, const uint8_t* __restrict__ pcb, const uint8_t* __restrict__ pcr, size_t pitch_y,
                        size_t pitch_c, int W, int H, int sub420, uint8_t* __restrict__ rgb, size_t pitch) {
    const size_t n = (size_t)W * H;
    const int cw = (W + 1) / 2, ch = (H + 1) / 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        const int Y = py[(size_t)y * pitch_y + x];
        int cb, cr;
        if (sub420) {
            cb = fancy_h2v2(pcb, pitch_c, cw, ch, x, y);
            cr = fancy_h2v2(pcr, pitch_c, cw, ch, x, y);
        } else {
            cb = pcb[(size_t)y * pitch_c + x];
            cr = pcr[(size_t)y * pitch_c + x];
        }
        ycc_to_rgb(Y, cb, cr, rgb + (size_t)y * pitch + 3 * (size_t)x);
    }
}

// ---------------------------------------------------------------- PSNR --
__global__ void __launch_bounds__(256) k_sq_err(const uint8_t* __restrict__ a, size_t pitch_a, const uint8_t* __restrict__ b, size_t pitch_b, int W,
                                                int H, unsigned long long* __restrict__ sum) {
    __shared__ unsigned long long s_part[8];
    const size_t n = (size_t)W * H;
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t y = i / W, x = i - y * W;
        const uint8_t *p = a + y * pitch_a + 3 * x, *q = b + y * pitch_b + 3 * x;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int d = (int)p[c] - (int)q[c];
            acc += (unsigned)(d * d);
        }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; ++i) t += s_part[i];
        atomicAdd(sum, t);
    }
}

// ---------------------------------------------------------------- launchers --
// d_cnt: chunks words, d_groups: chunks / 1024 + 1 words
int launch_rst_index(const uint8_t* d_scan, size_t n, uint32_t* d_cnt, uint32_t* d_groups, uint64_t* d_start, uint32_t max_int, cudaStream_t s) {
    const size_t chunks = (n + 255) / 256, groups = (chunks + 1023) / 1024;
    cudaMemsetAsync(d_start, 0, sizeof(uint64_t), s);  // interval 0 starts at the first byte of the scan
    if (!chunks) return 0;
    k_rst_count<<<(unsigned)((chunks + 127) / 128), 128, 0, s>>>(d_scan, n, d_cnt);
    k_rst_scan_groups<<<(unsigned)((groups + 63) / 64), 64, 0, s>>>(d_cnt, chunks, d_groups);
    k_rst_scan_totals<<<1, 1, 0, s>>>(d_groups, groups);
    k_rst_positions<<<(unsigned)((chunks + 127) / 128), 128, 0, s>>>(d_scan, n, d_cnt, d_groups, d_start, max_int);
    return 4;
}
int launch_huff_decode(const uint8_t* d_scan, size_t n, const uint64_t* d_start, uint32_t n_int, uint32_t ri, uint32_t n_mcu, int bpm,
                       const DecTables* d_tabs, const JfifInfo& info, int16_t* d_coef, cudaStream_t s) {
    cudaMemsetAsync(d_coef, 0, (size_t)n_mcu * bpm * 128, s);
    k_huff_decode<<<(n_int + 63) / 64, 64, 0, s>>>(d_scan, n, d_start, n_int, ri, n_mcu, bpm, d_tabs, info, d_coef);
    return 1;
}
int launch_reconstruct(const int16_t* d_coef, uint32_t n_mcu, int mcux, int bpm, const JfifInfo& info, uint8_t* py, uint8_t* pcb, uint8_t* pcr,
                       size_t pitch_y, size_t pitch_c, uint8_t* d_rgb, size_t pitch, cudaStream_t s) {
    const size_t nb = (size_t)n_mcu * bpm;
    k_idct<<<(unsigned)((nb + 127) / 128), 128, 0, s>>>(d_coef, n_mcu, mcux, bpm, info, py, pcb, pcr, pitch_y, pitch_c);
    const size_t n = (size_t)info.W * info.H, g = (n + 255) / 256;
    k_color<<<(unsigned)(g > 148 * 32 ? 148 * 32 : g), 256, 0, s>>>(py, pcb, pcr, pitch_y, pitch_c, (int)info.W, (int)info.H, info.sub == JB_SUB_420, d_rgb,
                                                                 pitch);
    return 2;
}
int launch_sq_err(const uint8_t* a, size_t pitch_a, const uint8_t* b, size_t pitch_b, size_t W, size_t H, unsigned long long* d_sum, cudaStream_t s) {
    cudaMemsetAsync(d_sum, 0, 8, s);
    const size_t n = W * H, g = (n + 255) / 256;
    k_sq_err<<<(unsigned)(g > 148 * 16 ? 148 * 16 : g), 256, 0, s>>>(a, pitch_a, b, pitch_b, (int)W, (int)H, d_sum);
    return 1;
}

}  // namespace jb
