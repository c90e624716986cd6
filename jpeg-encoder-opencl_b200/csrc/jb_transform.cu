// Fused per-MCU transform kernel for sm_100a:
//   RGB8 -> YCbCr (exact, utils.cpp:92-110) -> chroma 2x2 mean (utils.cpp:113-141)
//   -> mirror padding (utils.cpp:199-233) -> level shift (utils.cpp:190-196)
//   -> 8x8 FDCT (utils.cpp:314-347, true DCT) -> quantisation (utils.cpp:454-467)
//   -> zigzag (utils.cpp:539-558) -> int16 coefficients in scan order.
// plus the binary64 fix-up kernel that replays, in the reference's exact
// operation order, every coefficient whose binary32 quotient landed within the
// error band of a rounding tie, so that the stored coefficients equal the
// reference's (out-of-place) binary64 pipeline bit for bit.
//
// Work decomposition: one warp owns a "unit" (16 MCUs of 16x16 px in 4:2:0, 32
// MCUs of 8x8 px otherwise); one thread keeps a whole 8x8 block in registers, so
// both DCT passes, quantisation and the zigzag permutation need no shuffles.
// Results are staged through a per-warp, XOR-swizzled 4 KB shared-memory tile so
// that the HBM stores are coalesced 128-bit stores of whole 128-byte blocks.
// k_transform handles MCUs that lie completely inside the image; the few MCUs
// that need mirror padding (last MCU column/row of an image whose size is not a
// multiple of the MCU) go through k_transform_edge, a small generic kernel, which
// keeps the hot kernel free of the padding code (instruction-cache footprint).
// Algorithmic HBM traffic: 3 B/px read + 2 B/sample written = 6 B/px (4:2:0) or
// 9 B/px (4:4:4, replicated 4:2:0).
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_fp16.h>

#include "jb_pixels.cuh"

namespace jb {

__device__ __forceinline__ void fdct2d(float (&v)[64]) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
        fdct8(v[r * 8 + 0], v[r * 8 + 1], v[r * 8 + 2], v[r * 8 + 3], v[r * 8 + 4], v[r * 8 + 5], v[r * 8 + 6],
              v[r * 8 + 7]);
#pragma unroll
    for (int c = 0; c < 8; ++c)
        fdct8(v[c], v[8 + c], v[16 + c], v[24 + c], v[32 + c], v[40 + c], v[48 + c], v[56 + c]);
}

// DC coefficient in exact integer arithmetic.  v[0] is the exact integer sum S of
// the 64 level-shifted samples (the AAN DC path only adds), F = S/8, and the
// reference rounds S*c/q with c = fl(fl(1/sqrt2)^2/4) slightly below 1/8, i.e. an
// exact tie goes towards zero; the host checks this rule against the binary64
// expression for every S before enabling it (QuantConst::dc_exact).
template <int TAB>
__device__ __forceinline__ uint32_t quantize_dc(float s, const TransformArgs& a) {
    int S = __float2int_rn(s);
    uint32_t A = (uint32_t)abs(S);
    uint32_t m = __umulhi(2u * A + a.qc.dc_d[TAB] - 1u, a.qc.dc_m[TAB]);  // floor((2A + D - 1) / 2D)
    return (uint32_t)(S < 0 ? -(int)m : (int)m);
}

// Quantise the block in v (natural order), permute to zigzag, pack to int16 and
// write the 128 bytes to this lane's row of the warp's swizzled staging tile.
template <int TAB>
__device__ __forceinline__ void quant_stage(const float (&v)[64], const TransformArgs& a, uint4* st, int lane,
                                            uint32_t& tie_lo, uint32_t& tie_hi) {
    uint32_t wd[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int n0 = zz_nat(2 * j), n1 = zz_nat(2 * j + 1);
        bool t0, t1;
        uint32_t b0 = quantize_bits(v[n0], a.qc.mul[TAB][n0], a.qc.band[TAB][n0], t0);
        uint32_t b1 = quantize_bits(v[n1], a.qc.mul[TAB][n1], a.qc.band[TAB][n1], t1);
        if (j < 16) {
            if (t0) tie_lo |= 1u << (2 * j);
            if (t1) tie_lo |= 1u << (2 * j + 1);
        } else {
            if (t0) tie_hi |= 1u << (2 * j - 32);
            if (t1) tie_hi |= 1u << (2 * j - 31);
        }
        wd[j] = __byte_perm(b0, b1, 0x5410);
    }
    if (a.qc.dc_exact) {  // one uniform branch, outside the unrolled loop: DC in exact integer arithmetic
        wd[0] = __byte_perm(quantize_dc<TAB>(v[0], a), wd[0], 0x7610);
        tie_lo &= ~1u;
    }
#pragma unroll
    for (int p = 0; p < 8; ++p)
        st[lane * 8 + (p ^ (lane & 7))] = make_uint4(wd[4 * p], wd[4 * p + 1], wd[4 * p + 2], wd[4 * p + 3]);
}

// The flagged coefficients of one block take consecutive places of the list (one atomic per block, not per
// coefficient): k_fixup then fetches the block's samples once for all of them.
__device__ __noinline__ void append_ties(uint32_t* list, uint32_t* count, uint32_t cap, uint32_t gblock, uint32_t lo,
                                         uint32_t hi) {
    uint32_t idx = atomicAdd(count, (uint32_t)(__popc(lo) + __popc(hi)));
    while (lo | hi) {
        int k;
        if (lo) {
            k = __ffs(lo) - 1;
            lo &= lo - 1;
        } else {
            k = 31 + __ffs(hi);
            hi &= hi - 1;
        }
        if (idx < cap) list[idx] = gblock * 64u + (uint32_t)k;
        ++idx;
    }
}

// Copy the 32 staged blocks of the warp to HBM with 128-bit stores.  In a 6-block
// MCU lanes (2m, 2m+1) hold blocks (blk_base, blk_base+1) of MCU m; in a 3-block
// MCU lane m holds block blk_base of MCU m.
template <int BPM>
__device__ __forceinline__ void copy_out(const uint4* st, uint4* coef4, size_t mcu_g0, int mcus_valid, int blk_base,
                                         int lane) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        int g = it * 32 + lane, sb = g >> 3, p = g & 7;
        int mcu = BPM == 6 ? sb >> 1 : sb;
        int blk = blk_base + (BPM == 6 ? (sb & 1) : 0);
        if (mcu < mcus_valid) coef4[((mcu_g0 + mcu) * BPM + blk) * 8 + p] = st[sb * 8 + (p ^ (sb & 7))];
    }
}

// ---- colour conversion of one row of 8 pixels (24 bytes in w) ----------------
// T values are the 8.24 fixed-point numbers of jb_math.h; the three multiply-adds
// per channel are written as PTX so that they stay three IMADs.
__device__ __forceinline__ uint32_t mad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t byte_of(uint32_t w, int k) { return __byte_perm(w, 0, 0x4440 + k); }

struct Row8 {
    int y[8];         // level-shifted luma
    uint32_t cb[8];   // T_cb, T_cr: the chroma byte is the top byte
    uint32_t cr[8];
};

// SHARED: `ydown` is the copy of the tie table in shared memory and the eight look-ups of a row with a
// tie are issued unconditionally (branch-free); otherwise predicated loads from the global table (the
// 4:2:0 CUDA-core kernel has no registers to spare for the former).
template <bool SHARED>
__device__ __forceinline__ void csc_row8(const uint32_t (&w)[6], const uint32_t* __restrict__ ydown, Row8& o) {
    uint32_t ty[8], tmin = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = 3 * i;
        uint32_t R = byte_of(w[k >> 2], k & 3), G = byte_of(w[(k + 1) >> 2], (k + 1) & 3),
                 B = byte_of(w[(k + 2) >> 2], (k + 2) & 3);
        ty[i] = mad(B, KY_B, mad(G, KY_G, mad(R, KY_R, 0x80000000u)));  // T_y - 128.0
        o.cb[i] = mad(B, 0x00800000u, mad(G, 0u - KCB_G, mad(R, 0u - KCB_R, 0x80000000u)));
        o.cr[i] = mad(R, 0x00800000u, mad(B, 0u - KCR_B, mad(G, 0u - KCR_G, 0x80000000u)));
        tmin = min(tmin, ty[i] & Y_TIE_MASK);
        o.y[i] = (int)ty[i] >> 24;
    }
    if (tmin == 0) {  // some pixel of the row is a CSC tie (exact integer luma): consult the table
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = 3 * i;
            if (SHARED) {
                uint32_t idx = (byte_of(w[k >> 2], k & 3) << 8) | byte_of(w[(k + 1) >> 2], (k + 1) & 3);
                uint32_t bit = (ydown[idx >> 5] >> (idx & 31)) & 1u;
                o.y[i] -= (int)((ty[i] & Y_TIE_MASK) == 0 ? bit : 0u);
            } else if ((ty[i] & Y_TIE_MASK) == 0) {
                uint32_t idx = (byte_of(w[k >> 2], k & 3) << 8) | byte_of(w[(k + 1) >> 2], (k + 1) & 3);
                o.y[i] -= (int)((__ldg(ydown + (idx >> 5)) >> (idx & 31)) & 1u);
            }
        }
    }
}

// ------------------------------------------------------------------ 4:2:0 --
// Lane = one 8-pixel-wide, 16-row half of a 16x16 MCU: two luma blocks, then one
// chroma block (left lane Cb, right lane Cr).  Chroma samples are parked as
// floats in shared memory: word address m*136 + ch*68 + row*8 + col (the pads
// make both the 128-bit writes and the 128-bit reads bank-conflict free).
constexpr int CH_MCU_STRIDE = 136, CH_BLK_STRIDE = 68, CH_WARP_WORDS = 16 * CH_MCU_STRIDE;

template <int ALIGN>
__device__ __forceinline__ void unit_420(const TransformArgs& a, const Image& im, size_t mcu_g0, int mcu_x0,
                                         int mcus_valid, int my, uint4* st, float* ch, int lane) {
    const int half = lane & 1;
    const bool valid = (lane >> 1) < mcus_valid;
    // lanes without an MCU of their own redo the last valid one (no stores): the code stays
    // convergent, which lets the CTA-wide barriers sit anywhere
    const int m = valid ? lane >> 1 : max(mcus_valid - 1, 0);
    const uint32_t gb0 = (uint32_t)(mcu_g0 + m) * 6u;
    uint4* coef4 = reinterpret_cast<uint4*>(a.coef);
    float v[64];
    const uint8_t* col0 = im.base + (size_t)((mcu_x0 + m) * 16 + half * 8) * 3;

#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
        __syncthreads();  // keep the CTA's warps in step (instruction cache), see TW
        {
            // rows below an (even) image height are mirrored (utils.cpp:223-232); with H even the mirrored
            // row pair is again a complete 2x2-cell pair, so the chroma means are the reference's
            uint32_t w[8][6];
#pragma unroll
            for (int r = 0; r < 8; ++r)
                load24<ALIGN>(col0 + (size_t)mirror(my * 16 + h * 8 + r, im.H) * im.pitch, w[r]);
            uint32_t pb[4], pr[4];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                Row8 o;
                csc_row8<false>(w[r], im.ydown, o);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[r * 8 + i] = (float)o.y[i];
                uint32_t sb[4], sr[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    sb[c] = (o.cb[2 * c] >> 24) + (o.cb[2 * c + 1] >> 24);
                    sr[c] = (o.cr[2 * c] >> 24) + (o.cr[2 * c + 1] >> 24);
                }
                if (r & 1) {  // utils.cpp:126-127: truncated mean of the 2x2 cell, then level shift
                    float4 fb, fr;
                    fb.x = (float)((int)((pb[0] + sb[0]) >> 2) - 128);
                    fb.y = (float)((int)((pb[1] + sb[1]) >> 2) - 128);
                    fb.z = (float)((int)((pb[2] + sb[2]) >> 2) - 128);
                    fb.w = (float)((int)((pb[3] + sb[3]) >> 2) - 128);
                    fr.x = (float)((int)((pr[0] + sr[0]) >> 2) - 128);
                    fr.y = (float)((int)((pr[1] + sr[1]) >> 2) - 128);
                    fr.z = (float)((int)((pr[2] + sr[2]) >> 2) - 128);
                    fr.w = (float)((int)((pr[3] + sr[3]) >> 2) - 128);
                    float* dst = ch + m * CH_MCU_STRIDE + (h * 4 + (r >> 1)) * 8 + half * 4;
                    *reinterpret_cast<float4*>(dst) = fb;
                    *reinterpret_cast<float4*>(dst + CH_BLK_STRIDE) = fr;
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        pb[c] = sb[c];
                        pr[c] = sr[c];
                    }
                }
            }
            fdct2d(v);
            uint32_t tl = 0, th = 0;
            quant_stage<0>(v, a, st, lane, tl, th);
            if (valid && (tl | th)) append_ties(a.tie_list, a.tie_count, a.tie_cap, gb0 + 2 * h + half, tl, th);
        }
        __syncwarp();
        copy_out<6>(st, coef4, mcu_g0, mcus_valid, 2 * h, lane);
        __syncwarp();
    }
    __syncthreads();
    {
        const float* src = ch + m * CH_MCU_STRIDE + half * CH_BLK_STRIDE;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float4 f = *reinterpret_cast<const float4*>(src + j * 4);
            v[4 * j] = f.x;
            v[4 * j + 1] = f.y;
            v[4 * j + 2] = f.z;
            v[4 * j + 3] = f.w;
        }
        fdct2d(v);
        uint32_t tl = 0, th = 0;
        quant_stage<1>(v, a, st, lane, tl, th);
        if (valid && (tl | th)) append_ties(a.tie_list, a.tie_count, a.tie_cap, gb0 + 4 + half, tl, th);
    }
    __syncwarp();
    copy_out<6>(st, coef4, mcu_g0, mcus_valid, 4, lane);
    __syncwarp();
}

// ------------------------------------------------- 4:4:4 / replicated 4:2:0 --
// Lane = one 8x8 MCU: Y, Cb, Cr blocks in turn; chroma is parked as packed bytes.
__device__ __forceinline__ uint32_t pack_s8(uint32_t acc, int val, int pos) {
    return acc | (((uint32_t)val & 0xFFu) << (8 * pos));
}
__device__ __forceinline__ float unpack_s8(uint32_t w, int pos) { return (float)(int)(int8_t)(w >> (8 * pos)); }

template <int ALIGN, bool CDS>
__device__ __forceinline__ void unit_444(const TransformArgs& a, const Image& im, size_t mcu_g0, int mcu_x0,
                                         int mcus_valid, int my, uint4* st, int lane) {
    const bool valid = lane < mcus_valid;
    const int ml = valid ? lane : max(mcus_valid - 1, 0);  // convergent code: see unit_420
    const uint32_t gb0 = (uint32_t)(mcu_g0 + ml) * 3u;
    uint4* coef4 = reinterpret_cast<uint4*>(a.coef);
    float v[64];
    uint32_t cq[2][8][2];
    {
        const uint8_t* col0 = im.base + (size_t)((mcu_x0 + ml) * 8) * 3;
        uint32_t pb[4], pr[4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint32_t w[6];
            load24<ALIGN>(col0 + (size_t)mirror(my * 8 + r, im.H) * im.pitch, w);  // mirrored below the image
            Row8 o;
            csc_row8<true>(w, im.ydown, o);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[r * 8 + i] = (float)o.y[i];
            if (!CDS) {
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    uint32_t qb = 0, qr = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        qb = pack_s8(qb, (int)(o.cb[4 * k + i] >> 24) - 128, i);
                        qr = pack_s8(qr, (int)(o.cr[4 * k + i] >> 24) - 128, i);
                    }
                    cq[0][r][k] = qb;
                    cq[1][r][k] = qr;
                }
            } else {
                uint32_t sb[4], sr[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    sb[c] = (o.cb[2 * c] >> 24) + (o.cb[2 * c + 1] >> 24);
                    sr[c] = (o.cr[2 * c] >> 24) + (o.cr[2 * c + 1] >> 24);
                }
                if (r & 1) {  // the truncated 2x2 mean is written back to all four pixels (utils.cpp:130-138)
                    uint32_t qb[2] = {0, 0}, qr[2] = {0, 0};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        int mb = (int)((pb[c] + sb[c]) >> 2) - 128, mr = (int)((pr[c] + sr[c]) >> 2) - 128;
                        qb[c >> 1] = pack_s8(pack_s8(qb[c >> 1], mb, (2 * c) & 3), mb, (2 * c + 1) & 3);
                        qr[c >> 1] = pack_s8(pack_s8(qr[c >> 1], mr, (2 * c) & 3), mr, (2 * c + 1) & 3);
                    }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        cq[0][r - 1][k] = cq[0][r][k] = qb[k];
                        cq[1][r - 1][k] = cq[1][r][k] = qr[k];
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        pb[c] = sb[c];
                        pr[c] = sr[c];
                    }
                }
            }
        }
        fdct2d(v);
        uint32_t tl = 0, th = 0;
        quant_stage<0>(v, a, st, lane, tl, th);
        if (valid && (tl | th)) append_ties(a.tie_list, a.tie_count, a.tie_cap, gb0, tl, th);
    }
    __syncwarp();
    copy_out<3>(st, coef4, mcu_g0, mcus_valid, 0, lane);
    __syncwarp();
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
        __syncthreads();
        {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[r * 8 + i] = unpack_s8(c ? cq[1][r][i >> 2] : cq[0][r][i >> 2], i & 3);
            fdct2d(v);
            uint32_t tl = 0, th = 0;
            quant_stage<1>(v, a, st, lane, tl, th);
            if (valid && (tl | th)) append_ties(a.tie_list, a.tie_count, a.tie_cap, gb0 + 1 + c, tl, th);
        }
        __syncwarp();
        copy_out<3>(st, coef4, mcu_g0, mcus_valid, 1 + c, lane);
        __syncwarp();
    }
}

// TW warps per CTA.  One CTA of 16 warps per SM, with CTA-wide barriers at the stage boundaries,
// keeps all resident warps within a short window of the (long, fully unrolled) instruction
// stream, so that they share instruction-cache fetches instead of thrashing it.
constexpr int TW = 16;

template <int SUB, int ALIGN>
__global__ void __launch_bounds__(TW * 32, 16 / TW) k_transform(const __grid_constant__ TransformArgs a) {
    // dynamic shared memory: 4 staging tiles of 4 KB, then (4:2:0 only) 4 chroma parking areas
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_ydown[SUB == JB_SUB_420 ? 1 : 2048];  // 8x8 MCUs: the CSC tie table (jb_math.h) on chip
    if (SUB != JB_SUB_420) {
        for (int i = threadIdx.x; i < 2048; i += TW * 32) s_ydown[i] = __ldg(a.ydown + i);
        __syncthreads();
    }
    float* chroma = reinterpret_cast<float*>(smem + TW * 256);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* st = smem + warp * 256;
    const uint32_t units_per_frame = (uint32_t)a.units_per_row * (uint32_t)a.fast_mcuy;
    const int mcus_per_unit = SUB == JB_SUB_420 ? 16 : 32;
    for (uint32_t base = blockIdx.x * TW; base < a.total_units; base += gridDim.x * TW) {
        const bool active = base + warp < a.total_units;  // uniform trip count: every warp reaches the barriers
        const uint32_t unit = active ? base + warp : base;
        uint32_t f = unit / units_per_frame, rem = unit - f * units_per_frame;
        int my = (int)(rem / (uint32_t)a.units_per_row), ux = (int)(rem - (uint32_t)my * (uint32_t)a.units_per_row);
        Image im{a.rgb + (size_t)f * a.frame_stride, a.pitch, a.g.W, a.g.H, SUB == JB_SUB_420 ? a.ydown : s_ydown};
        int mcu_x0 = ux * mcus_per_unit;
        int mcus_valid = active ? min(mcus_per_unit, a.fast_mcux - mcu_x0) : 0;
        size_t mcu_g0 = (size_t)f * (size_t)a.g.n_mcu + (size_t)my * (size_t)a.g.mcux + (size_t)mcu_x0;
        if (SUB == JB_SUB_420)
            unit_420<ALIGN>(a, im, mcu_g0, mcu_x0, mcus_valid, my, st, chroma + warp * CH_WARP_WORDS, lane);
        else
            unit_444<ALIGN, SUB == JB_SUB_REPL420>(a, im, mcu_g0, mcu_x0, mcus_valid, my, st, lane);
    }
}

// ---------------------------------------------------------------- edge MCUs --
// One thread per block of an MCU that reaches into the mirror padding.  Samples
// come from ycc_at (the reference's CSC -> CDS -> pad order at any coordinate);
// the DCT / quantisation arithmetic is the same binary32 code as the hot kernel.
__global__ void __launch_bounds__(64) k_transform_edge(const __grid_constant__ TransformArgs a) {
    const int bpm = a.g.bpm;
    const uint32_t ncol = (uint32_t)(a.g.mcux - a.fast_mcux);                     // trailing MCU columns (0, 1 or 2)
    const uint32_t col_mcus = ncol * (uint32_t)a.g.mcuy;
    const uint32_t row_mcus = a.fast_mcuy < a.g.mcuy ? (uint32_t)a.fast_mcux : 0u; // the last MCU row (minus the corner)
    const uint32_t per_frame = (col_mcus + row_mcus) * (uint32_t)bpm;
    const uint32_t total = per_frame * (uint32_t)a.n_frames;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        uint32_t f = t / per_frame, r = t - f * per_frame;
        uint32_t e = r / (uint32_t)bpm, blk = r - e * (uint32_t)bpm;
        int mx, my;
        if (e < col_mcus) {
            my = (int)(e / ncol);
            mx = a.fast_mcux + (int)(e - (uint32_t)my * ncol);
        } else {
            mx = (int)(e - col_mcus);
            my = a.g.mcuy - 1;
        }
        int comp, x0, y0, step;
        if (a.g.sub == JB_SUB_420) {
            comp = blk < 4 ? 0 : (int)blk - 3;
            x0 = mx * 16 + (blk < 4 ? (int)(blk & 1) * 8 : 0);
            y0 = my * 16 + (blk < 4 ? (int)(blk >> 1) * 8 : 0);
            step = blk < 4 ? 1 : 2;  // chroma sample (i,j) = replicated plane at (2i,2j)
        } else {
            comp = (int)blk;
            x0 = mx * 8;
            y0 = my * 8;
            step = 1;
        }
        Image im{a.rgb + (size_t)f * a.frame_stride, a.pitch, a.g.W, a.g.H, a.ydown};
        float v[64];
        for (int y = 0; y < 8; ++y)
            for (int x = 0; x < 8; ++x) {
                uint32_t Y, Cb, Cr;
                if (a.g.sub == JB_SUB_444)
                    ycc_at<false>(im, x0 + x * step, y0 + y * step, Y, Cb, Cr);
                else
                    ycc_at<true>(im, x0 + x * step, y0 + y * step, Y, Cb, Cr);
                v[y * 8 + x] = (float)((int)(comp == 0 ? Y : comp == 1 ? Cb : Cr) - 128);
            }
        for (int i = 0; i < 8; ++i)
            fdct8(v[i * 8], v[i * 8 + 1], v[i * 8 + 2], v[i * 8 + 3], v[i * 8 + 4], v[i * 8 + 5], v[i * 8 + 6], v[i * 8 + 7]);
        for (int i = 0; i < 8; ++i)
            fdct8(v[i], v[8 + i], v[16 + i], v[24 + i], v[32 + i], v[40 + i], v[48 + i], v[56 + i]);
        const int tab = comp ? 1 : 0;
        uint32_t gblock = (uint32_t)(((size_t)f * a.g.n_mcu + (size_t)my * a.g.mcux + mx) * bpm + blk);
        int16_t* out = a.coef + (size_t)gblock * 64;
        for (int k = 0; k < 64; ++k) {
            int n = c_zz[k];
            bool tie = false;
            uint32_t bits;
            if (a.inplace_dct) {  // Q1: this kernel's AAN transform is the true DCT -- the replay computes the block
                bits = 0;
                tie = true;
            } else if (k == 0 && a.qc.dc_exact)
                bits = tab ? quantize_dc<1>(v[0], a) : quantize_dc<0>(v[0], a);
            else
                bits = quantize_bits(v[n], a.qc.mul[tab][n], a.qc.band[tab][n], tie);
            out[k] = (int16_t)(bits & 0xFFFFu);
            if (tie) {
                uint32_t idx = atomicAdd(a.tie_count, 1u);
                if (idx < a.tie_cap) a.tie_list[idx] = gblock * 64u + (uint32_t)k;
            }
        }
    }
}

// ---------------------------------------------------------------- NV12-style input --
// The image arrives as YCbCr 4:2:0 already (a Y plane and a half-resolution plane of interleaved Cb,Cr pairs, as video
// decoders and camera pipelines leave it in HBM): no colour conversion, no chroma averaging -- the reference's stages
// from the mirror padding on (utils.cpp:199-233, 190-196, 314-347, 454-467, 539-558).  Chroma sample (i, j) of an MCU
// row is the replicated full-resolution plane at the mirrored coordinate of (2i, 2j), exactly as in the RGB path.
// One thread per 8x8 block, binary32 AAN transform in registers, analytic near-tie band, binary64 replay by k_fixup.
__device__ __forceinline__ uint32_t nv12_sample(const uint8_t* yp, size_t pitch_y, const uint8_t* uvp, size_t pitch_uv, int W, int H, int x,
                                                int y, int comp) {  // (x, y): padded full-resolution coordinate
    const int sx = mirror(x, W), sy = mirror(y, H);
    return comp == 0 ? __ldg(yp + (size_t)sy * pitch_y + sx) : __ldg(uvp + (size_t)(sy >> 1) * pitch_uv + (size_t)(sx >> 1) * 2 + (comp - 1));
}

template <int TAB>
__device__ __forceinline__ void quant_global(const float (&v)[64], const TransformArgs& a, uint4* dst, uint32_t& tie_lo, uint32_t& tie_hi) {
    uint32_t wd[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int n0 = zz_nat(2 * j), n1 = zz_nat(2 * j + 1);
        bool t0, t1;
        uint32_t b0 = quantize_bits(v[n0], a.qc.mul[TAB][n0], a.qc.band[TAB][n0], t0);
        uint32_t b1 = quantize_bits(v[n1], a.qc.mul[TAB][n1], a.qc.band[TAB][n1], t1);
        if (j < 16) {
            if (t0) tie_lo |= 1u << (2 * j);
            if (t1) tie_lo |= 1u << (2 * j + 1);
        } else {
            if (t0) tie_hi |= 1u << (2 * j - 32);
            if (t1) tie_hi |= 1u << (2 * j - 31);
        }
        wd[j] = __byte_perm(b0, b1, 0x5410);
    }
    if (a.qc.dc_exact) {
        wd[0] = __byte_perm(quantize_dc<TAB>(v[0], a), wd[0], 0x7610);
        tie_lo &= ~1u;
    }
#pragma unroll
    for (int p = 0; p < 8; ++p) dst[p] = make_uint4(wd[4 * p], wd[4 * p + 1], wd[4 * p + 2], wd[4 * p + 3]);
}

__global__ void __launch_bounds__(128) k_transform_nv12(const __grid_constant__ TransformArgs a) {
    const uint32_t per_frame = (uint32_t)a.g.n_mcu * 6u, total = per_frame * (uint32_t)a.n_frames;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        const uint32_t f = t / per_frame, r = t - f * per_frame, mcu = r / 6u, blk = r - mcu * 6u;
        const int my = (int)(mcu / (uint32_t)a.g.mcux), mx = (int)(mcu - (uint32_t)my * (uint32_t)a.g.mcux);
        if (mx < a.fast_mcux && my < a.fast_mcuy) continue;  // k_transform_tc_nv12's part (fast_* = 0 without it)
        const uint8_t* yp = a.rgb + (size_t)f * a.frame_stride;
        const uint8_t* uvp = a.uv + (size_t)f * a.frame_stride_uv;
        float v[64];
        if (blk < 4) {
            const int x0 = mx * 16 + (int)(blk & 1) * 8, y0 = my * 16 + (int)(blk >> 1) * 8;
            const bool inside = x0 + 8 <= a.g.W;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint8_t* row = yp + (size_t)mirror(y0 + j, a.g.H) * a.pitch;
#pragma unroll
                for (int i = 0; i < 8; ++i) v[j * 8 + i] = (float)((int)__ldg(row + (inside ? x0 + i : mirror(x0 + i, a.g.W))) - 128);
            }
        } else {
            const int comp = (int)blk - 3;
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    v[j * 8 + i] = (float)((int)nv12_sample(yp, a.pitch, uvp, a.pitch_uv, a.g.W, a.g.H, 2 * (mx * 8 + i), 2 * (my * 8 + j), comp) - 128);
        }
        fdct2d(v);
        uint32_t tl = 0, th = 0;
        uint4* dst = reinterpret_cast<uint4*>(a.coef) + (size_t)t * 8;
        if (blk < 4)
            quant_global<0>(v, a, dst, tl, th);
        else
            quant_global<1>(v, a, dst, tl, th);
        if (tl | th) append_ties(a.tie_list, a.tie_count, a.tie_cap, t, tl, th);
    }
}

// RGB8 -> NV12 by the reference's own arithmetic: performCSC (utils.cpp:92-110, exact) and performCDS (utils.cpp:113-141):
// a complete 2x2 cell carries the truncated mean of its four chroma values, a cell cut by an odd edge the chroma of its
// top-left pixel.  One thread per cell.
__global__ void k_rgb_to_nv12(const uint8_t* __restrict__ rgb, int W, int H, size_t pitch, const uint32_t* __restrict__ ydown,
                              uint8_t* __restrict__ yo, size_t pitch_y, uint8_t* __restrict__ uvo, size_t pitch_uv) {
    const int cw = (W + 1) / 2, ch = (H + 1) / 2;
    const size_t n = (size_t)cw * ch;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
        const int cx = (int)(c % cw), cy = (int)(c / cw), x = 2 * cx, y = 2 * cy;
        const bool whole = x + 1 < W && y + 1 < H;
        uint32_t scb = 0, scr = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                if (x + i >= W || y + j >= H) continue;
                const uint8_t* p = rgb + (size_t)(y + j) * pitch + (size_t)(x + i) * 3;
                const uint32_t r = p[0], g = p[1], b = p[2];
                yo[(size_t)(y + j) * pitch_y + x + i] = (uint8_t)csc_y(r, g, b, ydown);
                if (whole || (i == 0 && j == 0)) {
                    scb += csc_cb(r, g, b);
                    scr += csc_cr(r, g, b);
                }
            }
        uvo[(size_t)cy * pitch_uv + 2 * cx] = (uint8_t)(whole ? scb >> 2 : scb);
        uvo[(size_t)cy * pitch_uv + 2 * cx + 1] = (uint8_t)(whole ? scr >> 2 : scr);
    }
}

int launch_rgb_to_nv12(const uint8_t* rgb, size_t W, size_t H, size_t pitch, const uint32_t* ydown, uint8_t* y, size_t pitch_y, uint8_t* uv,
                       size_t pitch_uv, cudaStream_t s) {
    const size_t n = ((W + 1) / 2) * ((H + 1) / 2), g = (n + 255) / 256;
    k_rgb_to_nv12<<<(int)(g > 148 * 32 ? 148 * 32 : g), 256, 0, s>>>(rgb, (int)W, (int)H, pitch, ydown, y, pitch_y, uv, pitch_uv);
    return 1;
}

// ================================================================ tensor-core variant ==
// 4:2:0 only.  The block transform (FDCT + quantiser scale + zigzag) is one 64x64 contraction
// per block: D[128 blocks][64 coefficients] = A[128 blocks][64 samples] x W^T on the 5th-gen
// tensor cores (tcgen05.mma, kind::f16, fp16 operands, fp32 accumulators in TMEM).  The 8-bit
// level-shifted samples are exact in fp16; W (times 2^10) is split into hi+lo fp16 matrices (22
// significand bits), so one tile costs 2 x 4 MMAs of M128 N64 K16.  A group of 128 threads owns
// one M=128 tile: thread i writes the 64 samples of its block as row i of the A tile (one
// 128-bit store per image row, 128-byte swizzle), so no register array of samples exists and
// the row loop stays rolled (small code: no instruction-cache pressure).  After the MMAs the
// thread reads row i of D back (tcgen05.ld 32x32b.x64): its own 64 scaled coefficients, already
// in zigzag order, for rounding, near-tie flagging and packing as in the FMA kernel.
constexpr int TC_GROUPS = 4;                       // 128-thread groups (M tiles) per CTA, one CTA per SM
constexpr int TC_TMEM_COLS = 512;                  // 2 accumulators x 64 columns per group, power of two
constexpr int TC_B_BYTES = 4 * 8192;               // [table][split] 64x64 fp16
constexpr int TC_TILE_BYTES = 128 * 128;           // one A tile
constexpr int TC_ROW_BYTES = 16 * 16 * 3;          // one image row of a warp's unit (16 MCUs of RGB)
constexpr int TC_RING_BYTES = 2 * 2 * 4 * TC_ROW_BYTES;  // raw-pixel ring of a group: [slot][row of the pair][warp]
constexpr int TC_SMEM = TC_B_BYTES + TC_GROUPS * (2 * TC_TILE_BYTES + TC_RING_BYTES) + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, 8-row groups 1024 B apart
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc),
        "r"(accumulate));
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mbar), "r"(parity)
            : "memory");
}

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, "
        "%46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ uint32_t bf16x2(int lo, int hi) {  // two small integers -> packed fp16 (exact)
    __half2 v = __floats2half2_rn((float)lo, (float)hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// Colour conversion of one row of 8 pixels for the tensor-core kernel: the three T values (8.24
// fixed point, jb_math.h) per pixel, luma already corrected on CSC ties (top byte = Y).
struct Row8T {
    uint32_t y[8], cb[8], cr[8];
};
__device__ __forceinline__ void csc_row8_t(const uint32_t (&w)[6], const uint32_t* ydown_sh, Row8T& o) {
    uint32_t tmin = 0xFFFFFFFFu;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = 3 * i;
        uint32_t R = byte_of(w[k >> 2], k & 3), G = byte_of(w[(k + 1) >> 2], (k + 1) & 3),
                 B = byte_of(w[(k + 2) >> 2], (k + 2) & 3);
        o.y[i] = mad(B, KY_B, mad(G, KY_G, R * KY_R));
        o.cb[i] = mad(B, 0x00800000u, mad(G, 0u - KCB_G, mad(R, 0u - KCB_R, 0x80000000u)));
        o.cr[i] = mad(R, 0x00800000u, mad(B, 0u - KCR_B, mad(G, 0u - KCR_G, 0x80000000u)));
        tmin = min(tmin, o.y[i] & Y_TIE_MASK);
    }
    if (tmin == 0) {  // some pixel of the row is a CSC tie (exact integer luma): consult the table
        // (8 KB, in static shared memory: no global round trip; all eight look-ups are issued unconditionally
        // so that the path stays branch-free, and only the tied pixels take the correction)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = 3 * i;
            const uint32_t R = byte_of(w[k >> 2], k & 3), G = byte_of(w[(k + 1) >> 2], (k + 1) & 3);
            // word R*8 + G/32 (one IMAD + one shift), bit G % 32 (the funnel shift takes G mod 32 itself)
            const uint32_t word = ydown_sh[mad(R, 256u, G) >> 5];
            uint32_t y = o.y[i];
            // the fraction of a tie is tiny: "top byte - 1" is a subtraction of 2^24
            asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                "shf.r.wrap.b32 t, %1, %1, %2;\n\t"
                "and.b32 t, t, 1;\n\t"
                "setp.ne.u32 p, t, 0;\n\t"
                "setp.eq.and.u32 p, %3, 0, p;\n\t"
                "@p sub.u32 %0, %0, 0x1000000;\n\t}"
                : "+r"(y)
                : "r"(word), "r"(G), "r"(y & Y_TIE_MASK));
            o.y[i] = y;
        }
    }
}
// fp16 pair (Y0 - 128, Y1 - 128) from the top bytes of two T values: 0x6400 | n is the fp16 number
// 1024 + n, and subtracting 1152 is exact.
__device__ __forceinline__ uint32_t and_or(uint32_t x, uint32_t m, uint32_t c) {  // (x & m) | c as ONE LOP3
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(x), "r"(m), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t luma_h2(uint32_t t0, uint32_t t1) {
    uint32_t bits = and_or(__byte_perm(t0, t1, 0x7773), 0x00FF00FFu, 0x64006400u);
    __half2 h = __hsub2(*reinterpret_cast<__half2*>(&bits), __floats2half2_rn(1152.0f, 1152.0f));
    return *reinterpret_cast<uint32_t*>(&h);
}
// fp16 pair (floor(s0/4) - 128, floor(s1/4) - 128) from two sums of four chroma bytes (s < 1024):
// 0x6400 | (s & ~3) is 1024 + 4 floor(s/4); one exact FMA does the rest (utils.cpp:126-127).
__device__ __forceinline__ uint32_t chroma_h2(uint32_t s0, uint32_t s1) {
    uint32_t bits = and_or(__byte_perm(s0, s1, 0x5410), 0x03FC03FCu, 0x64006400u);
    __half2 h = __hfma2(*reinterpret_cast<__half2*>(&bits), __floats2half2_rn(0.25f, 0.25f),
                        __floats2half2_rn(-384.0f, -384.0f));
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    return ((uint64_t)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t sub_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// Round the 64 scaled coefficients t (zigzag order) of one block, flag near ties, pack, stage.
template <int TAB>
__device__ __forceinline__ void tc_quant_stage(const uint32_t (&t)[64], const TransformArgs& a, uint4* st, int lane,
                                               uint32_t& tie_lo, uint32_t& tie_hi, uint32_t wait_mbar, uint32_t wait_parity) {
    uint32_t wd[32];
    const float inv = (float)(1.0 / JB_TC_W_SCALE);
    // two coefficients per instruction (packed fp32 FFMA2 / FADD2 of sm_100): r = x inv + 1.5 2^23 rounds to
    // the nearest integer, d = x inv - (r - 1.5 2^23) is the distance from it
    const uint64_t inv2 = pack_f32x2(inv, inv), magic2 = pack_f32x2(JB_ROUND_MAGIC, JB_ROUND_MAGIC);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const uint64_t x2 = ((uint64_t)t[2 * j + 1] << 32) | t[2 * j];
        const uint64_t r2 = fma_f32x2(x2, inv2, magic2);
        const uint64_t d2 = fma_f32x2(x2, inv2, sub_f32x2(magic2, r2));
        if (fabsf(__uint_as_float((uint32_t)d2)) > a.tband[TAB][2 * j]) {
            if (j < 16) tie_lo |= 1u << (2 * j); else tie_hi |= 1u << (2 * j - 32);
        }
        if (fabsf(__uint_as_float((uint32_t)(d2 >> 32))) > a.tband[TAB][2 * j + 1]) {
            if (j < 16) tie_lo |= 1u << (2 * j + 1); else tie_hi |= 1u << (2 * j - 31);
        }
        wd[j] = __byte_perm((uint32_t)r2, (uint32_t)(r2 >> 32), 0x5410);
    }
    if (a.qc.dc_exact) {  // one uniform branch, outside the unrolled loop
        // t[0] ~ 2^10 S/(8q): recover the exact integer sample sum S, then the integer DC rule
        uint32_t dc = quantize_dc<TAB>(__uint_as_float(t[0]) * (inv * (float)a.qc.dc_d[TAB]), a);
        wd[0] = __byte_perm(dc, wd[0], 0x7610);
        tie_lo &= ~1u;
    }
    if (wait_mbar) mbar_wait(wait_mbar, wait_parity);  // the staging area is still an MMA operand until then
#pragma unroll
    for (int p = 0; p < 8; ++p)
        st[lane * 8 + (p ^ (lane & 7))] = make_uint4(wd[4 * p], wd[4 * p + 1], wd[4 * p + 2], wd[4 * p + 3]);
}

__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
    uint32_t p;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ uint32_t lds_volatile(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const uint8_t* src) {
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst), "l"(src), "n"(BYTES) : "memory");
}

// One warp's unit: a run of 16 consecutive MCUs of the kernel's MCU sequence (frame-major, then MCU rows,
// then MCU columns of the region the kernel covers).  A run may wrap to the next MCU row or frame, so every
// lane pair keeps the coordinates of its own MCU.
struct TcUnit {
    const uint8_t* ptr;  // first byte of the MCU in image row 0 (frame base + 48 bytes per MCU column)
    int y0;              // first image row of the MCU
    uint32_t gm;         // index of the MCU in the coefficient array (frame-major over all MCUs of a frame)
    bool valid;          // the MCU exists
};

template <int ALIGN>
__global__ void __launch_bounds__(TC_GROUPS * 128, 1) k_transform_tc(const __grid_constant__ TransformArgs a) {
    extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar[TC_GROUPS][2];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_next[TC_GROUPS];
    __shared__ uint32_t s_desc[TC_GROUPS + 1][4];  // descriptor low words: [group]{luma tile, chroma tile}, [TC_GROUPS][table * 2 + split]
    __shared__ uint32_t s_ydown[2048];  // the CSC tie table (jb_math.h), 8 KB
    // keep the address arithmetic on the shared-space pointer (1024-byte alignment for the 128B swizzle)
    uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, g = tid >> 7, gt = tid & 127, wg = gt >> 5, lane = tid & 31;
    uint8_t* sB = smem;
    uint8_t* tileA = smem + TC_B_BYTES + g * 2 * TC_TILE_BYTES;  // luma rows 0-7, then 8-15; staging between
    uint8_t* tileC = tileA + TC_TILE_BYTES;                      // chroma
    // ring of raw pixels: address of this warp's row `r` (0/1) of slot `sl` (0/1)
    const uint32_t ring = smem_u32(smem + TC_B_BYTES + TC_GROUPS * 2 * TC_TILE_BYTES + g * TC_RING_BYTES) + wg * TC_ROW_BYTES;
    // ---- one-time setup: W matrices, tensor memory, barriers ---------------------------------------
    for (int i = tid; i < TC_B_BYTES / 16; i += TC_GROUPS * 128)
        reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(a.tc_mat) + i);
    for (int i = tid; i < 2048; i += TC_GROUPS * 128) s_ydown[i] = __ldg(a.ydown + i);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "n"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (gt == 0) {
        s_desc[g][0] = (uint32_t)umma_desc(smem_u32(tileA));
        s_desc[g][1] = (uint32_t)umma_desc(smem_u32(tileC));
        s_desc[TC_GROUPS][g] = (uint32_t)umma_desc(smem_u32(sB + g * 8192));
    }
    if (tid < TC_GROUPS * 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar[0][0]) + 8 * tid));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d0 = s_tmem + (uint32_t)(g * 128), tmem_d1 = tmem_d0 + 64;  // two accumulators per group
    const uint32_t lane_off = (uint32_t)(wg * 32) << 16;
    const uint32_t mbar0 = smem_u32(&s_mbar[g][0]), mbar1 = smem_u32(&s_mbar[g][1]);
    const uint32_t idesc = (1u << 4) | (8u << 17) | (8u << 24);  // f32 += fp16 x fp16 (formats 0), N=64, M=128
    uint32_t phase0 = 0, phase1 = 0;
    uint4* coef4 = reinterpret_cast<uint4*>(a.coef);
    const int half = lane & 1;
    const uint32_t stride = gridDim.x * TC_GROUPS * 4;
    // shared addresses of this thread's A rows: own row (luma), and the rows that take its chroma
    const uint32_t sw_own = (uint32_t)(gt & 7);
    const uint32_t a_row = smem_u32(tileA) + gt * 128;
    const int row_cb = gt & ~1, row_cr = gt | 1;
    const uint32_t ac_cb = smem_u32(tileC) + row_cb * 128 + half * 8, ac_cr = smem_u32(tileC) + row_cr * 128 + half * 8;
    const uint32_t sw_cb = (uint32_t)(row_cb & 7), sw_cr = (uint32_t)(row_cr & 7);

    // the 8 MMAs of one tile, issued by one elected lane of the group's first warp; completion arrives on
    // `mbar`.  Every operand is computed from warp-uniform values (the group index comes out of a shuffle, so
    // the compiler keeps it in a uniform register): no per-thread descriptor arithmetic, and no
    // serialisation loop around each tcgen05.mma.
    const bool issuer = __shfl_sync(0xffffffffu, (uint32_t)wg, 0) == 0;
    auto issue = [&](int tile_sel, int tab, uint32_t tmem_d, uint32_t mbar) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
            const uint64_t hi = (uint64_t)0x40004040u << 32;
            const uint64_t da = hi | lds_volatile(smem_u32(&s_desc[g][tile_sel]));
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                const uint64_t db = hi | lds_volatile(smem_u32(&s_desc[TC_GROUPS][tab * 2 + s2]));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (s2 | k) ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
        }
    };
    // writes of this thread to the tiles become visible to the tensor core, all threads of the group arrive
    // (a full group barrier: arrive/wait for the non-issuing warps and CTA-wide barriers both measured slower)
    auto publish = [&]() {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
    };
    // n / d and n % d with the host's m = floor(2^32 / d): the estimate is at most one too small
    auto divmod = [](uint32_t n, uint32_t d, uint32_t m, uint32_t& q, uint32_t& r) {
        q = __umulhi(n, m);
        r = n - q * d;
        if (r >= d) {
            q += 1;
            r -= d;
        }
    };
    // every warp of the group runs the same control flow; a warp past the end gets an empty unit
    auto decode = [&](uint32_t unit_base) {
        const uint32_t lin = (unit_base + wg) * 16u + (uint32_t)(lane >> 1);  // this lane pair's MCU
        TcUnit u;
        u.valid = unit_base + wg < a.total_units && lin < a.tc_mcus;
        const uint32_t l = u.valid ? lin : 0u;
        uint32_t f, rem, my, mx;
        divmod(l, a.tc_per_frame, a.tc_magic_frame, f, rem);
        divmod(rem, a.tc_row_len, a.tc_magic_row, my, mx);
        u.ptr = a.rgb + (size_t)f * a.frame_stride + (size_t)mx * 48;
        u.y0 = (int)my * 16;
        u.gm = f * (uint32_t)a.g.n_mcu + my * (uint32_t)a.g.mcux + mx;
        return u;
    };
    // fetch cursor: where this lane's share of the unit's rows comes from.  16-byte aligned images are copied
    // warp-cooperatively: chunk c (16 bytes) of the 48 chunks of a row belongs to MCU c / 3 — lane takes chunks
    // lane and, for lane < 16, 32 + lane, and gets the coordinates of their MCUs from the owning lane pairs.
    // Otherwise every lane copies the 24 bytes of its own half MCU.
    const uint32_t lane_share = (uint32_t)lane * (ALIGN == 16 ? 16u : 24u);
    const uint32_t ring_wr = ring + lane_share, ring_rd = ring + (uint32_t)lane * 24u;
    // Each cursor walks down the image two rows per fetch: `p` = first row of the next pair, `y` = its
    // (unmirrored) row index.  Rows at or below the image height are mirrored (utils.cpp:211-233): the hot
    // kernel only meets them with an even height (plan_fast), so a row pair is either inside (step +pitch)
    // or mirrored (H-1, H-2, ...: step -pitch), and the walk turns around between two pairs.
    const int pitch_i = (int)a.pitch, img_h = a.g.H;
    const uint8_t *fc_p0 = nullptr, *fc_p1 = nullptr;
    int fc_y0 = 0, fc_y1 = 0;
    bool fc_v0 = false, fc_v1 = false;
    auto shfl_ptr = [&](const uint8_t* p, int src) {
        unsigned long long v = (unsigned long long)p;
        uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
        return (const uint8_t*)(((unsigned long long)hi << 32) | lo);
    };
    auto aim = [&](const TcUnit& u) {
        const uint8_t* row0 = u.ptr + (uint64_t)(uint32_t)u.y0 * (uint32_t)pitch_i;  // y0 < H: never mirrored
        if (ALIGN == 16) {
            const int c0 = lane, c1 = 32 + (lane & 15);
            const int m0 = c0 / 3, m1 = c1 / 3;
            fc_p0 = shfl_ptr(row0, 2 * m0) + (c0 - 3 * m0) * 16;
            fc_y0 = __shfl_sync(0xffffffffu, u.y0, 2 * m0);
            fc_v0 = __shfl_sync(0xffffffffu, (int)u.valid, 2 * m0) != 0;
            fc_p1 = shfl_ptr(row0, 2 * m1) + (c1 - 3 * m1) * 16;
            fc_y1 = __shfl_sync(0xffffffffu, u.y0, 2 * m1);
            fc_v1 = __shfl_sync(0xffffffffu, (int)u.valid, 2 * m1) != 0 && lane < 16;
        } else {
            fc_p0 = row0 + half * 24;
            fc_y0 = u.y0;
            fc_v0 = u.valid;
        }
    };
    // the two rows of a cursor's next pair, then advance the cursor
    auto walk = [&](const uint8_t*& p, int& y, const uint8_t*& r0, const uint8_t*& r1) {
        const int d = y < img_h ? pitch_i : -pitch_i;
        r0 = p;
        r1 = p + d;
        y += 2;
        p = r1 + (y == img_h ? 0 : d);
    };
    // start the copy of the next row pair of every MCU of the unit into ring slot `sl`: one commit group
    auto fetch_pair = [&](int sl) {
        const uint32_t d0 = ring_wr + (uint32_t)sl * (2 * 4 * TC_ROW_BYTES), d1 = d0 + 4 * TC_ROW_BYTES;
        const uint8_t *r0, *r1;
        walk(fc_p0, fc_y0, r0, r1);
        if (ALIGN == 16) {
            if (fc_v0) {
                cp_async<16>(d0, r0);
                cp_async<16>(d1, r1);
            }
            walk(fc_p1, fc_y1, r0, r1);
            if (fc_v1) {
                cp_async<16>(d0 + 512, r0);
                cp_async<16>(d1 + 512, r1);
            }
        } else if (fc_v0) {
#pragma unroll
            for (int j = 0; j < 24 / ALIGN; ++j) {
                cp_async<ALIGN>(d0 + j * ALIGN, r0 + j * ALIGN);
                cp_async<ALIGN>(d1 + j * ALIGN, r1 + j * ALIGN);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // Strips are handed out dynamically (the first one statically): groups that meet cheaper content or
    // emptier edge strips take more of them, so that all SMs finish together.  One thread of the group
    // draws the next index at the start of a unit; the group reads it after the first barrier of the unit.
    uint32_t base = (blockIdx.x * TC_GROUPS + g) * 4;
    TcUnit cur = decode(base);
    if (base < a.total_units) {
        aim(cur);
        fetch_pair(0);
        fetch_pair(1);
    }
    while (base < a.total_units) {
        if (gt == 0) s_next[g] = stride + atomicAdd(a.unit_counter, 4u);
        uint32_t nbase = 0;
        TcUnit nxt = cur;
        const bool valid = cur.valid;
        const uint32_t gm = valid ? cur.gm : 0xFFFFFFFFu;  // all ones: no such MCU
        const uint32_t gb0 = cur.gm * 6u;

        // ---- 16 image rows = 8 row pairs: ring -> registers -> colour conversion -> A tiles -----------
        // The ring holds two pairs; the refill of a slot (the pair after next, possibly of the next unit)
        // is issued right after the slot has been read, two pairs of work ahead of its use.
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            if (ALIGN == 16) __syncwarp();  // lanes read bytes that other lanes copied
            const uint32_t src = ring_rd + (uint32_t)(it & 1) * (2 * 4 * TC_ROW_BYTES);
            uint32_t w0[6], w1[6];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                uint2 v0 = lds64(src + 8 * j), v1 = lds64(src + 4 * TC_ROW_BYTES + 8 * j);
                w0[2 * j] = v0.x;
                w0[2 * j + 1] = v0.y;
                w1[2 * j] = v1.x;
                w1[2 * j + 1] = v1.y;
            }
            if (ALIGN == 16) __syncwarp();  // every lane has read the slot before anyone refills it
            if (it == 6) {
                nbase = s_next[g];  // written before the barrier of it == 3
                nxt = decode(nbase);  // past the end: an empty unit, nothing is fetched
                aim(nxt);
            }
            fetch_pair(it & 1);
            if (it == 4) mbar_wait(mbar0, phase0);  // the MMAs of rows 0-7 have consumed the tile: overwrite it

            Row8T o0, o1;
            csc_row8_t(w0, s_ydown, o0);
            csc_row8_t(w1, s_ydown, o1);
            const uint32_t rp = (uint32_t)(it & 3);
            sts128(a_row + (((2 * rp) ^ sw_own) << 4), make_uint4(luma_h2(o0.y[0], o0.y[1]), luma_h2(o0.y[2], o0.y[3]),
                                                                  luma_h2(o0.y[4], o0.y[5]), luma_h2(o0.y[6], o0.y[7])));
            sts128(a_row + (((2 * rp + 1) ^ sw_own) << 4), make_uint4(luma_h2(o1.y[0], o1.y[1]), luma_h2(o1.y[2], o1.y[3]),
                                                                      luma_h2(o1.y[4], o1.y[5]), luma_h2(o1.y[6], o1.y[7])));
            uint32_t sb[4], sr[4];  // sums of the four chroma bytes of each 2x2 cell
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                sb[c] = (o0.cb[2 * c] >> 24) + (o0.cb[2 * c + 1] >> 24) + (o1.cb[2 * c] >> 24) + (o1.cb[2 * c + 1] >> 24);
                sr[c] = (o0.cr[2 * c] >> 24) + (o0.cr[2 * c + 1] >> 24) + (o1.cr[2 * c] >> 24) + (o1.cr[2 * c + 1] >> 24);
            }
            const uint32_t crow = (uint32_t)it;  // chroma row = K chunk of the chroma tile
            sts64(ac_cb + ((crow ^ sw_cb) << 4), make_uint2(chroma_h2(sb[0], sb[1]), chroma_h2(sb[2], sb[3])));
            sts64(ac_cr + ((crow ^ sw_cr) << 4), make_uint2(chroma_h2(sr[0], sr[1]), chroma_h2(sr[2], sr[3])));
            if (it == 3) {
                publish();
                if (issuer) issue(0, 0, tmem_d0, mbar0);
            }
        }
        publish();
        if (issuer) issue(0, 0, tmem_d1, mbar1);

        // read this thread's row of the accumulator, round / flag / pack, stage in tileA, store blocks blk, blk+1
        auto finish = [&](uint32_t tmem_d, int tab, int blk, uint32_t wait_mbar, uint32_t wait_parity) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t t[64];
            tmem_ld64(tmem_d + lane_off, t);
            uint4* st = reinterpret_cast<uint4*>(tileA) + wg * 256;  // this warp's 32 rows of the tile
            uint32_t tl = 0, th = 0;
            if (tab == 0)
                tc_quant_stage<0>(t, a, st, lane, tl, th, wait_mbar, wait_parity);
            else
                tc_quant_stage<1>(t, a, st, lane, tl, th, wait_mbar, wait_parity);
            if (valid && (tl | th)) append_ties(a.tie_list, a.tie_count, a.tie_cap, gb0 + blk + half, tl, th);
            __syncwarp();
            // coalesced copy-out: 8 x 32 pieces of 16 bytes; piece g belongs to block g >> 3 of the warp, i.e. to
            // MCU g >> 4, whose index sits in lane pair g >> 4 (32-bit piece indices: n_blocks * 8 < 2^32)
            const uint32_t piece0 = (uint32_t)(blk + ((lane >> 3) & 1)) * 8u + (uint32_t)(lane & 7);
#pragma unroll
            for (int it8 = 0; it8 < 8; ++it8) {
                const int sb = it8 * 4 + (lane >> 3), pc = lane & 7;
                const uint32_t m_gm = __shfl_sync(0xffffffffu, gm, 4 * it8 + 2 * (lane >> 4));
                if (m_gm != 0xFFFFFFFFu) coef4[m_gm * 48u + piece0] = st[sb * 8 + (pc ^ (sb & 7))];
            }
            __syncwarp();
        };
        phase0 ^= 1;                                  // (rows 0-7: completion already observed at it == 4)
        finish(tmem_d0, 0, 0, mbar1, phase1);         // Y00 / Y01; staging waits for the MMAs of rows 8-15
        phase1 ^= 1;
        publish();                                    // every thread has read accumulator 0: it takes the chroma tile
        if (issuer) issue(1, 1, tmem_d0, mbar0);
        finish(tmem_d1, 0, 2, 0, 0);                  // Y10 / Y11
        mbar_wait(mbar0, phase0);
        phase0 ^= 1;
        finish(tmem_d0, 1, 4, 0, 0);                  // Cb / Cr
        cur = nxt;
        base = nbase;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(TC_TMEM_COLS));
}

// ================================================== tensor-core transform for NV12-style input (4:2:0) ==========
// k_transform_tc with the colour conversion taken out: the frames are a Y plane and a plane of interleaved Cb,Cr
// pairs already, so a row pair of a unit is 2 x 256 luma bytes + 256 chroma bytes (16 bytes per MCU each), copied
// warp-cooperatively into the ring (lanes 0-15: luma row 0 and the chroma row, lanes 16-31: luma row 1), and a
// sample becomes its fp16 operand with one byte permute (0x6400 | byte = 1024 + byte) and half an exact HADD2 (-1152).
// Everything after the A tiles -- MMA schedule, epilogue, near-tie list -- is k_transform_tc's.  Covers the complete
// MCU columns (16-byte aligned planes, pitches and frame strides); k_transform_nv12 takes a partial last column.
struct TcUnitNv {
    const uint8_t *ptr, *puv;  // first byte of the MCU in luma row 0 / chroma row 0 of the frame
    int y0;
    uint32_t gm;
    bool valid;
};
__global__ void __launch_bounds__(TC_GROUPS * 128, 1) k_transform_tc_nv12(const __grid_constant__ TransformArgs a) {
    extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar[TC_GROUPS][2];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_next[TC_GROUPS];
    __shared__ uint32_t s_desc[TC_GROUPS + 1][4];  // descriptor low words: [group]{luma tile, chroma tile}, [TC_GROUPS][table * 2 + split]
    // keep the address arithmetic on the shared-space pointer (1024-byte alignment for the 128B swizzle)
    uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, g = tid >> 7, gt = tid & 127, wg = gt >> 5, lane = tid & 31;
    uint8_t* sB = smem;
    uint8_t* tileA = smem + TC_B_BYTES + g * 2 * TC_TILE_BYTES;  // luma rows 0-7, then 8-15; staging between
    uint8_t* tileC = tileA + TC_TILE_BYTES;                      // chroma
    // ring of raw pixels: address of this warp's row `r` (luma 0, luma 1, chroma) of slot `sl` (0-3)
    constexpr uint32_t NV_ROW = 256, NV_ROWS = 4 * NV_ROW, NV_SLOT = 3 * NV_ROWS;  // per group: [slot][luma 0, luma 1, chroma][warp]
    static_assert(4 * NV_SLOT <= TC_RING_BYTES, "the NV12 ring (four slots: three row pairs ahead) fits the RGB ring");
    const uint32_t ring = smem_u32(smem + TC_B_BYTES + TC_GROUPS * 2 * TC_TILE_BYTES + g * TC_RING_BYTES) + wg * NV_ROW;
    // ---- one-time setup: W matrices, tensor memory, barriers ---------------------------------------
    for (int i = tid; i < TC_B_BYTES / 16; i += TC_GROUPS * 128)
        reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(a.tc_mat) + i);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "n"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (gt == 0) {
        s_desc[g][0] = (uint32_t)umma_desc(smem_u32(tileA));
        s_desc[g][1] = (uint32_t)umma_desc(smem_u32(tileC));
        s_desc[TC_GROUPS][g] = (uint32_t)umma_desc(smem_u32(sB + g * 8192));
    }
    if (tid < TC_GROUPS * 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar[0][0]) + 8 * tid));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d0 = s_tmem + (uint32_t)(g * 128), tmem_d1 = tmem_d0 + 64;  // two accumulators per group
    const uint32_t lane_off = (uint32_t)(wg * 32) << 16;
    const uint32_t mbar0 = smem_u32(&s_mbar[g][0]), mbar1 = smem_u32(&s_mbar[g][1]);
    const uint32_t idesc = (1u << 4) | (8u << 17) | (8u << 24);  // f32 += fp16 x fp16 (formats 0), N=64, M=128
    uint32_t phase0 = 0, phase1 = 0;
    uint4* coef4 = reinterpret_cast<uint4*>(a.coef);
    const int half = lane & 1;
    const uint32_t stride = gridDim.x * TC_GROUPS * 4;
    // shared addresses of this thread's A rows: own row (luma), and the rows that take its chroma
    const uint32_t sw_own = (uint32_t)(gt & 7);
    const uint32_t a_row = smem_u32(tileA) + gt * 128;
    const int row_cb = gt & ~1, row_cr = gt | 1;
    const uint32_t ac_cb = smem_u32(tileC) + row_cb * 128 + half * 8, ac_cr = smem_u32(tileC) + row_cr * 128 + half * 8;
    const uint32_t sw_cb = (uint32_t)(row_cb & 7), sw_cr = (uint32_t)(row_cr & 7);

    // the 8 MMAs of one tile, issued by one elected lane of the group's first warp; completion arrives on
    // `mbar`.  Every operand is computed from warp-uniform values (the group index comes out of a shuffle, so
    // the compiler keeps it in a uniform register): no per-thread descriptor arithmetic, and no
    // serialisation loop around each tcgen05.mma.
    const bool issuer = __shfl_sync(0xffffffffu, (uint32_t)wg, 0) == 0;
    auto issue = [&](int tile_sel, int tab, uint32_t tmem_d, uint32_t mbar) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
            const uint64_t hi = (uint64_t)0x40004040u << 32;
            const uint64_t da = hi | lds_volatile(smem_u32(&s_desc[g][tile_sel]));
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                const uint64_t db = hi | lds_volatile(smem_u32(&s_desc[TC_GROUPS][tab * 2 + s2]));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (s2 | k) ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
        }
    };
    // writes of this thread to the tiles become visible to the tensor core, all threads of the group arrive
    // (a full group barrier: arrive/wait for the non-issuing warps and CTA-wide barriers both measured slower)
    auto publish = [&]() {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
    };
    // n / d and n % d with the host's m = floor(2^32 / d): the estimate is at most one too small
    auto divmod = [](uint32_t n, uint32_t d, uint32_t m, uint32_t& q, uint32_t& r) {
        q = __umulhi(n, m);
        r = n - q * d;
        if (r >= d) {
            q += 1;
            r -= d;
        }
    };
    // every warp of the group runs the same control flow; a warp past the end gets an empty unit
    auto decode = [&](uint32_t unit_base) {
        const uint32_t lin = (unit_base + wg) * 16u + (uint32_t)(lane >> 1);  // this lane pair's MCU
        TcUnitNv u;
        u.valid = unit_base + wg < a.total_units && lin < a.tc_mcus;
        const uint32_t l = u.valid ? lin : 0u;
        uint32_t f, rem, my, mx;
        divmod(l, a.tc_per_frame, a.tc_magic_frame, f, rem);
        divmod(rem, a.tc_row_len, a.tc_magic_row, my, mx);
        u.ptr = a.rgb + (size_t)f * a.frame_stride + (size_t)mx * 16;
        u.puv = a.uv + (size_t)f * a.frame_stride_uv + (size_t)mx * 16;
        u.y0 = (int)my * 16;
        u.gm = f * (uint32_t)a.g.n_mcu + my * (uint32_t)a.g.mcux + mx;
        return u;
    };
    // fetch cursors: chunk c (16 bytes) of a row belongs to MCU c of the unit.  Lane L copies chunk L & 15 of luma row
    // L >> 4 of the pair, lanes 0-15 also chunk L of the chroma row; the coordinates come from the owning lane pairs.
    // A cursor is the MCU's column in the frame plus a row counter: rows at or below the image height are mirrored
    // (utils.cpp:211-233: padded row y reads row 2H - y - 1; a chroma row is the row of the mirrored even luma row).
    const uint32_t ring_wr = ring + (uint32_t)(lane >> 4) * NV_ROWS + (uint32_t)(lane & 15) * 16u;
    const uint32_t ring_wr_c = ring + 2 * NV_ROWS + (uint32_t)(lane & 15) * 16u;
    const uint32_t ring_rd = ring + (uint32_t)lane * 8u;  // this lane's half MCU: 8 bytes of every row
    const int img_h = a.g.H;
    const uint8_t *fc_p0 = nullptr, *fc_p1 = nullptr;
    int fc_y = 0, fc_yc = 0;
    bool fc_v0 = false, fc_v1 = false;
    auto shfl_ptr = [&](const uint8_t* p, int src) {
        unsigned long long v = (unsigned long long)p;
        uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
        return (const uint8_t*)(((unsigned long long)hi << 32) | lo);
    };
    auto aim = [&](const TcUnitNv& u) {
        const int m = lane & 15, y0 = __shfl_sync(0xffffffffu, u.y0, 2 * m);
        fc_p0 = shfl_ptr(u.ptr, 2 * m);
        fc_p1 = shfl_ptr(u.puv, 2 * m);
        fc_y = y0 + (lane >> 4);
        fc_yc = y0;  // (the luma row whose chroma row is wanted)
        fc_v0 = __shfl_sync(0xffffffffu, (int)u.valid, 2 * m) != 0;
        fc_v1 = fc_v0 && lane < 16;
    };
    // start the copy of the next row pair (two luma rows, one chroma row) of every MCU of the unit into ring slot `sl`
    auto fetch_pair = [&](int sl) {
        if (fc_v0) cp_async<16>(ring_wr + (uint32_t)sl * NV_SLOT, fc_p0 + (size_t)(uint32_t)mirror(fc_y, img_h) * a.pitch);
        if (fc_v1) cp_async<16>(ring_wr_c + (uint32_t)sl * NV_SLOT, fc_p1 + (size_t)((uint32_t)mirror(fc_yc, img_h) >> 1) * a.pitch_uv);
        fc_y += 2;
        fc_yc += 2;
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // Strips are handed out dynamically (the first one statically): groups that meet cheaper content or
    // emptier edge strips take more of them, so that all SMs finish together.  One thread of the group
    // draws the next index at the start of a unit; the group reads it after the first barrier of the unit.
    uint32_t base = (blockIdx.x * TC_GROUPS + g) * 4;
    TcUnitNv cur = decode(base);
    if (base < a.total_units) {
        aim(cur);
        fetch_pair(0);
        fetch_pair(1);
        fetch_pair(2);
    }
    while (base < a.total_units) {
        if (gt == 0) s_next[g] = stride + atomicAdd(a.unit_counter, 4u);
        uint32_t nbase = 0;
        TcUnitNv nxt = cur;
        const bool valid = cur.valid;
        const uint32_t gm = valid ? cur.gm : 0xFFFFFFFFu;  // all ones: no such MCU
        const uint32_t gb0 = cur.gm * 6u;

        // ---- 16 image rows = 8 row pairs: ring -> registers -> colour conversion -> A tiles -----------
        // The ring holds four pairs; the refill of a slot (three pairs on, possibly of the next unit) is issued
        // one iteration after the slot has been read, three pairs of work ahead of its use.
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            asm volatile("cp.async.wait_group 2;" ::: "memory");
            __syncwarp();  // lanes read bytes that other lanes copied
            const uint32_t src = ring_rd + (uint32_t)(it & 3) * NV_SLOT;
            const uint2 l0 = lds64(src), l1 = lds64(src + NV_ROWS), cc = lds64(src + 2 * NV_ROWS);
            __syncwarp();  // every lane has read the slot before anyone refills it
            if (it == 5) {  // (the ring runs three pairs ahead: pairs 8, 9, 10 are the next unit's first three)
                nbase = s_next[g];  // written before the barrier of it == 3
                nxt = decode(nbase);  // past the end: an empty unit, nothing is fetched
                aim(nxt);
            }
            fetch_pair((it + 3) & 3);  // the slot read in the previous iteration
            if (it == 4) {
                mbar_wait(mbar0, phase0);  // the MMAs of rows 0-7 have consumed the tile: overwrite it
                // With no colour conversion between them the two row pairs the ring runs ahead are ~300 cycles of work,
                // less than a trip to HBM: ask L2 for the rows of the NEXT unit now (its index was drawn at the top of
                // this one), a lane per row -- 16 luma rows, 8 chroma rows, 256 bytes each when the unit does not wrap.
                const TcUnitNv pf = decode(s_next[g]);
                const uint8_t *p0 = shfl_ptr(pf.ptr, 0), *p15 = shfl_ptr(pf.ptr, 30);
                const uint8_t* pc = shfl_ptr(pf.puv, 0);
                const int py = __shfl_sync(0xffffffffu, pf.y0, 0);
                // (the last MCU of the unit exists and sits 15 MCUs to the right of the first: no wrap, every address is a row's)
                if (__shfl_sync(0xffffffffu, (int)pf.valid, 30) && p15 == p0 + 15 * 16 && lane < 24) {
                    const uint8_t* r = lane < 16 ? p0 + (size_t)(uint32_t)mirror(py + lane, img_h) * a.pitch
                                                 : pc + (size_t)((uint32_t)mirror(py + 2 * (lane - 16), img_h) >> 1) * a.pitch_uv;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(r + 128));
                }
            }

            // bytes -> fp16 operands: 0x6400 | byte is 1024 + byte, minus 1152 is the level-shifted sample (exact)
            auto h2 = [](uint32_t w, uint32_t sel) {
                uint32_t bits = __byte_perm(w, 0x64646464u, sel);
                __half2 h = __hsub2(*reinterpret_cast<__half2*>(&bits), __floats2half2_rn(1152.0f, 1152.0f));
                return *reinterpret_cast<uint32_t*>(&h);
            };
            const uint32_t rp = (uint32_t)(it & 3);
            sts128(a_row + (((2 * rp) ^ sw_own) << 4), make_uint4(h2(l0.x, 0x4140), h2(l0.x, 0x4342), h2(l0.y, 0x4140), h2(l0.y, 0x4342)));
            sts128(a_row + (((2 * rp + 1) ^ sw_own) << 4), make_uint4(h2(l1.x, 0x4140), h2(l1.x, 0x4342), h2(l1.y, 0x4140), h2(l1.y, 0x4342)));
            const uint32_t crow = (uint32_t)it;  // chroma row = K chunk of the chroma tile
            sts64(ac_cb + ((crow ^ sw_cb) << 4), make_uint2(h2(cc.x, 0x4240), h2(cc.y, 0x4240)));  // Cb of four pairs
            sts64(ac_cr + ((crow ^ sw_cr) << 4), make_uint2(h2(cc.x, 0x4341), h2(cc.y, 0x4341)));  // Cr
            if (it == 3) {
                publish();
                if (issuer) issue(0, 0, tmem_d0, mbar0);
            }
        }
        publish();
        if (issuer) issue(0, 0, tmem_d1, mbar1);

        // read this thread's row of the accumulator, round / flag / pack, stage in tileA, store blocks blk, blk+1
        auto finish = [&](uint32_t tmem_d, int tab, int blk, uint32_t wait_mbar, uint32_t wait_parity) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t t[64];
            tmem_ld64(tmem_d + lane_off, t);
            uint4* st = reinterpret_cast<uint4*>(tileA) + wg * 256;  // this warp's 32 rows of the tile
            uint32_t tl = 0, th = 0;
            if (tab == 0)
                tc_quant_stage<0>(t, a, st, lane, tl, th, wait_mbar, wait_parity);
            else
                tc_quant_stage<1>(t, a, st, lane, tl, th, wait_mbar, wait_parity);
            if (valid && (tl | th)) append_ties(a.tie_list, a.tie_count, a.tie_cap, gb0 + blk + half, tl, th);
            __syncwarp();
            // coalesced copy-out: 8 x 32 pieces of 16 bytes; piece g belongs to block g >> 3 of the warp, i.e. to
            // MCU g >> 4, whose index sits in lane pair g >> 4 (32-bit piece indices: n_blocks * 8 < 2^32)
            const uint32_t piece0 = (uint32_t)(blk + ((lane >> 3) & 1)) * 8u + (uint32_t)(lane & 7);
#pragma unroll
            for (int it8 = 0; it8 < 8; ++it8) {
                const int sb = it8 * 4 + (lane >> 3), pc = lane & 7;
                const uint32_t m_gm = __shfl_sync(0xffffffffu, gm, 4 * it8 + 2 * (lane >> 4));
                if (m_gm != 0xFFFFFFFFu) coef4[m_gm * 48u + piece0] = st[sb * 8 + (pc ^ (sb & 7))];
            }
            __syncwarp();
        };
        phase0 ^= 1;                                  // (rows 0-7: completion already observed at it == 4)
        finish(tmem_d0, 0, 0, mbar1, phase1);         // Y00 / Y01; staging waits for the MMAs of rows 8-15
        phase1 ^= 1;
        publish();                                    // every thread has read accumulator 0: it takes the chroma tile
        if (issuer) issue(1, 1, tmem_d0, mbar0);
        finish(tmem_d1, 0, 2, 0, 0);                  // Y10 / Y11
        mbar_wait(mbar0, phase0);
        phase0 ^= 1;
        finish(tmem_d0, 1, 4, 0, 0);                  // Cb / Cr
        cur = nxt;
        base = nbase;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(TC_TMEM_COLS));
}

// ================================================== tensor-core variant with TMA-staged image tiles (4:2:0) ==
// Same pipeline as k_transform_tc; what changes is how pixels reach shared memory.  A warp's unit is a RECTANGLE of
// 8 x 2 MCUs (128 x 32 px), so every row pair of the unit is two boxes of a 3-D tensor map over the RGB batch
// (uint32 elements: x = pitch / 4, y = image row, z = frame; box 96 x 2 x 1 = 2 image rows x 384 bytes): one elected
// lane issues two cp.async.bulk.tensor (SASS UTMALDG) per row pair, completion arrives on an mbarrier.  The per-lane
// cp.async of k_transform_tc and its pointer-walking cursors (two 64-bit pointers, row counters, validity flags per
// lane; ~30 instructions per row pair in every lane) are gone; coordinates are warp-uniform integers.  Rows below an
// (even) image height are fetched from their mirror image (utils.cpp:223-232): the box of the pair (y, y+1) with
// y >= H is the box at 2H-2-y, and the lanes read its two rows in swapped order.  Needs a 16-byte aligned base,
// pitch and frame stride (the tensor map's rules); other inputs take k_transform_tc<8 / 4>.
constexpr int TM_BOX_BYTES = 2 * 8 * 48;                    // one box: 2 image rows x 8 MCUs x 48 bytes
constexpr int TM_SLOT_BYTES = 2 * TM_BOX_BYTES;             // one row pair of a warp's unit: MCU row 0, MCU row 1
constexpr int TM_RING_BYTES = 2 * 4 * TM_SLOT_BYTES;        // per group: [slot][warp]; = TC_RING_BYTES
static_assert(TM_RING_BYTES == TC_RING_BYTES, "the TMA ring takes the place of the cp.async ring");

__device__ __forceinline__ void tma_load_box(uint32_t dst, const void* tmap, int x, int y, int z, uint32_t mbar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(mbar)
        : "memory");
}

__global__ void __launch_bounds__(TC_GROUPS * 128, 1) k_transform_tma(const __grid_constant__ TransformArgs a,
                                                                      const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar[TC_GROUPS][2];
    __shared__ __align__(8) uint64_t s_ring_bar[TC_GROUPS][4][2];  // [group][warp][slot]: a row pair of the unit has landed
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_next[TC_GROUPS];
    __shared__ uint32_t s_desc[TC_GROUPS + 1][4];
    __shared__ uint32_t s_ydown[2048];
    uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, g = tid >> 7, gt = tid & 127, wg = gt >> 5, lane = tid & 31;
    uint8_t* sB = smem;
    uint8_t* tileA = smem + TC_B_BYTES + g * 2 * TC_TILE_BYTES;
    uint8_t* tileC = tileA + TC_TILE_BYTES;
    // ring of raw pixels: this warp's two slots (one row pair each: box of MCU row 0, box of MCU row 1)
    const uint32_t ring = smem_u32(smem + TC_B_BYTES + TC_GROUPS * 2 * TC_TILE_BYTES + g * TM_RING_BYTES) + wg * TM_SLOT_BYTES;
    constexpr uint32_t kSlotStride = 4 * TM_SLOT_BYTES;
    for (int i = tid; i < TC_B_BYTES / 16; i += TC_GROUPS * 128)
        reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(a.tc_mat) + i);
    for (int i = tid; i < 2048; i += TC_GROUPS * 128) s_ydown[i] = __ldg(a.ydown + i);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (gt == 0) {
        s_desc[g][0] = (uint32_t)umma_desc(smem_u32(tileA));
        s_desc[g][1] = (uint32_t)umma_desc(smem_u32(tileC));
        s_desc[TC_GROUPS][g] = (uint32_t)umma_desc(smem_u32(sB + g * 8192));
    }
    if (tid < TC_GROUPS * 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar[0][0]) + 8 * tid));
    if (tid < TC_GROUPS * 8) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_ring_bar[0][0][0]) + 8 * tid));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d0 = s_tmem + (uint32_t)(g * 128), tmem_d1 = tmem_d0 + 64;
    const uint32_t lane_off = (uint32_t)(wg * 32) << 16;
    const uint32_t mbar0 = smem_u32(&s_mbar[g][0]), mbar1 = smem_u32(&s_mbar[g][1]);
    const uint32_t rbar = smem_u32(&s_ring_bar[g][wg][0]);  // + 8 * slot
    const uint32_t idesc = (1u << 4) | (8u << 17) | (8u << 24);
    uint32_t phase0 = 0, phase1 = 0, rphase = 0;  // rphase: bit s = parity the next wait on ring slot s expects
    uint4* coef4 = reinterpret_cast<uint4*>(a.coef);
    const int half = lane & 1;
    const uint32_t stride = gridDim.x * TC_GROUPS * 4;
    const uint32_t sw_own = (uint32_t)(gt & 7);
    const uint32_t a_row = smem_u32(tileA) + gt * 128;
    const int row_cb = gt & ~1, row_cr = gt | 1;
    const uint32_t ac_cb = smem_u32(tileC) + row_cb * 128 + half * 8, ac_cr = smem_u32(tileC) + row_cr * 128 + half * 8;
    const uint32_t sw_cb = (uint32_t)(row_cb & 7), sw_cr = (uint32_t)(row_cr & 7);
    // this lane's MCU inside the unit: column c (0..7), MCU row r (0..1); its 24 bytes of a box row
    const int mc = (lane >> 1) & 7, mr = lane >> 4;
    const uint32_t lane_rd = ring + (uint32_t)mr * TM_BOX_BYTES + (uint32_t)mc * 48u + (uint32_t)half * 24u;
    const int img_h = a.g.H;

    const bool issuer = __shfl_sync(0xffffffffu, (uint32_t)wg, 0) == 0;
    auto issue = [&](int tile_sel, int tab, uint32_t tmem_d, uint32_t mbar) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
            const uint64_t hi = (uint64_t)0x40004040u << 32;
            const uint64_t da = hi | lds_volatile(smem_u32(&s_desc[g][tile_sel]));
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                const uint64_t db = hi | lds_volatile(smem_u32(&s_desc[TC_GROUPS][tab * 2 + s2]));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (s2 | k) ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
        }
    };
    auto publish = [&]() {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
    };
    auto divmod = [](uint32_t n, uint32_t d, uint32_t m, uint32_t& q, uint32_t& r) {
        q = __umulhi(n, m);
        r = n - q * d;
        if (r >= d) {
            q += 1;
            r -= d;
        }
    };
    // A warp's unit: warp-uniform coordinates (frame, first image row, first element column of its boxes)
    struct Unit {
        int f, y0, x_el;   // frame, first image row (MCU row 2 uy), first uint32 column (96 per 8 MCUs)
        bool valid, row1;  // the unit exists; its second MCU row exists
        uint32_t gm;       // this LANE's MCU in the coefficient array (0xFFFFFFFF: none)
    };
    auto decode = [&](uint32_t unit_base) {
        const uint32_t U = unit_base + (uint32_t)wg;
        Unit u;
        u.valid = U < a.total_units;
        uint32_t f, rem, uy, ux;
        divmod(u.valid ? U : 0u, a.tc_per_frame, a.tc_magic_frame, f, rem);
        divmod(rem, a.tc_row_len, a.tc_magic_row, uy, ux);
        u.f = (int)f;
        u.y0 = (int)uy * 32;
        u.x_el = (int)ux * 96;
        u.row1 = (int)(2 * uy + 1) < a.fast_mcuy;
        const int mx = (int)ux * 8 + mc, my = (int)uy * 2 + mr;
        u.gm = u.valid && mx < a.fast_mcux && my < a.fast_mcuy ? f * (uint32_t)a.g.n_mcu + (uint32_t)my * (uint32_t)a.g.mcux + (uint32_t)mx
                                                                : 0xFFFFFFFFu;
        return u;
    };
    // start the copy of row pair `pi` (0..7) of unit `u` into ring slot `sl`: lane 0 posts the byte count and the boxes
    auto fetch_pair = [&](const Unit& u, int pi, int sl) {
        if (lane == 0) {
            const uint32_t bar = rbar + 8u * (uint32_t)sl, dst = ring + (uint32_t)sl * kSlotStride;
            const uint32_t bytes = u.valid ? (u.row1 ? 2u * TM_BOX_BYTES : (uint32_t)TM_BOX_BYTES) : 0u;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            if (u.valid) {
                int ya = u.y0 + 2 * pi, yb = ya + 16;
                ya = ya < img_h ? ya : 2 * img_h - 2 - ya;  // mirrored pair: fetched from its mirror image, rows swapped by the readers
                yb = yb < img_h ? yb : 2 * img_h - 2 - yb;
                tma_load_box(dst, &tmap, u.x_el, ya, u.f, bar);
                if (u.row1) tma_load_box(dst + TM_BOX_BYTES, &tmap, u.x_el, yb, u.f, bar);
            }
        }
    };

    uint32_t base = (blockIdx.x * TC_GROUPS + g) * 4;
    Unit cur = decode(base);
    if (base < a.total_units) {
        fetch_pair(cur, 0, 0);
        fetch_pair(cur, 1, 1);
    }
    while (base < a.total_units) {
        if (gt == 0) s_next[g] = stride + atomicAdd(a.unit_counter, 4u);
        uint32_t nbase = 0;
        Unit nxt = cur;
        const uint32_t gm = cur.gm;
        const bool valid = gm != 0xFFFFFFFFu;
        const uint32_t gb0 = gm * 6u;
        const int y_lane = cur.y0 + 16 * mr;

#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            const int sl = it & 1;
            mbar_wait(rbar + 8u * (uint32_t)sl, (rphase >> sl) & 1u);
            rphase ^= 1u << sl;
            // the two image rows of this pair; below the image the box holds the mirror image: swapped order
            const bool swapped = y_lane + 2 * it >= img_h;
            const uint32_t src = lane_rd + (uint32_t)sl * kSlotStride;
            const uint32_t src0 = src + (swapped ? 384u : 0u), src1 = src + (swapped ? 0u : 384u);
            uint32_t w0[6], w1[6];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                uint2 v0 = lds64(src0 + 8 * j), v1 = lds64(src1 + 8 * j);
                w0[2 * j] = v0.x;
                w0[2 * j + 1] = v0.y;
                w1[2 * j] = v1.x;
                w1[2 * j + 1] = v1.y;
            }
            __syncwarp();  // every lane has read the slot before lane 0 lets the TMA refill it
            if (it == 6) {
                nbase = s_next[g];  // written before the barrier of it == 3
                nxt = decode(nbase);
            }
            if (it < 6)
                fetch_pair(cur, it + 2, sl);
            else
                fetch_pair(nxt, it - 6, sl);
            if (it == 4) mbar_wait(mbar0, phase0);

            Row8T o0, o1;
            csc_row8_t(w0, s_ydown, o0);
            csc_row8_t(w1, s_ydown, o1);
            const uint32_t rp = (uint32_t)(it & 3);
            sts128(a_row + (((2 * rp) ^ sw_own) << 4), make_uint4(luma_h2(o0.y[0], o0.y[1]), luma_h2(o0.y[2], o0.y[3]),
                                                                  luma_h2(o0.y[4], o0.y[5]), luma_h2(o0.y[6], o0.y[7])));
            sts128(a_row + (((2 * rp + 1) ^ sw_own) << 4), make_uint4(luma_h2(o1.y[0], o1.y[1]), luma_h2(o1.y[2], o1.y[3]),
                                                                      luma_h2(o1.y[4], o1.y[5]), luma_h2(o1.y[6], o1.y[7])));
            uint32_t sb[4], sr[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                sb[c] = (o0.cb[2 * c] >> 24) + (o0.cb[2 * c + 1] >> 24) + (o1.cb[2 * c] >> 24) + (o1.cb[2 * c + 1] >> 24);
                sr[c] = (o0.cr[2 * c] >> 24) + (o0.cr[2 * c + 1] >> 24) + (o1.cr[2 * c] >> 24) + (o1.cr[2 * c + 1] >> 24);
            }
            const uint32_t crow = (uint32_t)it;
            sts64(ac_cb + ((crow ^ sw_cb) << 4), make_uint2(chroma_h2(sb[0], sb[1]), chroma_h2(sb[2], sb[3])));
            sts64(ac_cr + ((crow ^ sw_cr) << 4), make_uint2(chroma_h2(sr[0], sr[1]), chroma_h2(sr[2], sr[3])));
            if (it == 3) {
                publish();
                if (issuer) issue(0, 0, tmem_d0, mbar0);
            }
        }
        publish();
        if (issuer) issue(0, 0, tmem_d1, mbar1);

        auto finish = [&](uint32_t tmem_d, int tab, int blk, uint32_t wait_mbar, uint32_t wait_parity) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t t[64];
            tmem_ld64(tmem_d + lane_off, t);
            uint4* st = reinterpret_cast<uint4*>(tileA) + wg * 256;
            uint32_t tl = 0, th = 0;
            if (tab == 0)
                tc_quant_stage<0>(t, a, st, lane, tl, th, wait_mbar, wait_parity);
            else
                tc_quant_stage<1>(t, a, st, lane, tl, th, wait_mbar, wait_parity);
            if (valid && (tl | th)) append_ties(a.tie_list, a.tie_count, a.tie_cap, gb0 + blk + half, tl, th);
            __syncwarp();
            const uint32_t piece0 = (uint32_t)(blk + ((lane >> 3) & 1)) * 8u + (uint32_t)(lane & 7);
#pragma unroll
            for (int it8 = 0; it8 < 8; ++it8) {
                const int sb = it8 * 4 + (lane >> 3), pc = lane & 7;
                const uint32_t m_gm = __shfl_sync(0xffffffffu, gm, 4 * it8 + 2 * (lane >> 4));
                if (m_gm != 0xFFFFFFFFu) coef4[m_gm * 48u + piece0] = st[sb * 8 + (pc ^ (sb & 7))];
            }
            __syncwarp();
        };
        phase0 ^= 1;
        finish(tmem_d0, 0, 0, mbar1, phase1);
        phase1 ^= 1;
        publish();
        if (issuer) issue(1, 1, tmem_d0, mbar0);
        finish(tmem_d1, 0, 2, 0, 0);
        mbar_wait(mbar0, phase0);
        phase0 ^= 1;
        finish(tmem_d0, 1, 4, 0, 0);
        cur = nxt;
        base = nbase;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(TC_TMEM_COLS));
}

// ================================================== tensor-core variant, 8x8-pixel MCUs ==
// 4:4:4 and replicated 4:2:0 (the reference's own mode, utils.cpp:113-141 + 667-695): an MCU is one 8x8
// block of each component.  Same machinery as k_transform_tc -- exact CSC on CUDA cores, one
// tcgen05 contraction per block, rounding / near-tie flagging / zigzag-ordered int16 stores in the epilogue --
// with these differences: a lane owns one MCU and a lane pair one *pair* of MCUs (48 contiguous bytes per image
// row, like a 16-pixel MCU), a unit is 16 consecutive pairs and 8 image rows; a group has three A tiles (Y,
// Cb, Cr: all three are produced by the same pass over the pixels), so only three groups fit in shared memory;
// the two accumulators take Y and Cb, then Cr once Y has been read.  The replicated 4:2:0 mode writes every
// 2x2 mean to its four samples of the chroma tiles.
// Replicated 4:2:0 has only 4x4 distinct chroma values per block (every 2x2 mean fills its cell), so its
// chroma contraction is over K = 16: D = C[128 blocks][16 means] x W'^T with W'[n][cell] the sum of the four W
// entries of the cell (built by the host in the chroma slots of the W buffer).  Cb and Cr means share one tile
// (32 bytes each per row), a group needs two tiles like the 4:2:0 kernel, and four groups fit again.
constexpr int t3_groups(int sub) { return sub == JB_SUB_REPL420 ? 4 : 3; }
constexpr int t3_tiles(int sub) { return sub == JB_SUB_REPL420 ? 2 : 3; }
constexpr int t3_smem(int sub) { return TC_B_BYTES + t3_groups(sub) * (t3_tiles(sub) * TC_TILE_BYTES + TC_RING_BYTES) + 1024; }

template <int SUB, int ALIGN>
__global__ void __launch_bounds__(t3_groups(SUB) * 128, 1) k_transform_tc3(const __grid_constant__ TransformArgs a) {
    constexpr bool REPL = SUB == JB_SUB_REPL420;
    constexpr int T3_GROUPS = t3_groups(SUB), T3_TILES = t3_tiles(SUB);
    extern __shared__ __align__(1024) uint8_t tc_smem_raw[];
    __shared__ __align__(8) uint64_t s_mbar[T3_GROUPS][2];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_next[T3_GROUPS], s_next0[T3_GROUPS];
    __shared__ uint32_t s_desc[T3_GROUPS + 1][4];  // descriptor low words: [group]{Y, Cb, Cr tile}, [T3_GROUPS][table * 2 + split]
    __shared__ uint32_t s_ydown[2048];             // the CSC tie table (jb_math.h), 8 KB
    uint8_t* smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, g = tid >> 7, gt = tid & 127, wg = gt >> 5, lane = tid & 31;
    uint8_t* sB = smem;
    uint8_t* tileY = smem + TC_B_BYTES + g * T3_TILES * TC_TILE_BYTES;  // Y, then Cb, then Cr (or one tile of chroma means); the Y tile doubles as staging
    const uint32_t ring = smem_u32(smem + TC_B_BYTES + T3_GROUPS * T3_TILES * TC_TILE_BYTES + g * TC_RING_BYTES) + wg * TC_ROW_BYTES;
    // ---- one-time setup: W matrices, tensor memory, barriers ---------------------------------------
    for (int i = tid; i < TC_B_BYTES / 16; i += T3_GROUPS * 128)
        reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(a.tc_mat) + i);
    for (int i = tid; i < 2048; i += T3_GROUPS * 128) s_ydown[i] = __ldg(a.ydown + i);
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)),
                     "n"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (gt < 3)  // replicated 4:2:0: Cr means start 32 bytes into the rows of the chroma tile
        s_desc[g][gt] = (uint32_t)umma_desc(smem_u32(tileY + (REPL && gt == 2 ? 1 : gt) * TC_TILE_BYTES)) + (REPL && gt == 2 ? 2u : 0u);
    if (tid < 4) s_desc[T3_GROUPS][tid] = (uint32_t)umma_desc(smem_u32(sB + tid * 8192));
    if (tid < T3_GROUPS * 2) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_mbar[0][0]) + 8 * tid));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d0 = s_tmem + (uint32_t)(g * 128), tmem_d1 = tmem_d0 + 64;  // two accumulators per group
    const uint32_t lane_off = (uint32_t)(wg * 32) << 16;
    const uint32_t mbar0 = smem_u32(&s_mbar[g][0]), mbar1 = smem_u32(&s_mbar[g][1]);
    const uint32_t idesc = (1u << 4) | (8u << 17) | (8u << 24);  // f32 += fp16 x fp16, N=64, M=128
    uint32_t phase0 = 0, phase1 = 0;
    uint4* coef4 = reinterpret_cast<uint4*>(a.coef);
    const int half = lane & 1;
    const uint32_t stride = gridDim.x * T3_GROUPS * 4;
    const uint32_t sw_own = (uint32_t)(gt & 7);
    const uint32_t y_row = smem_u32(tileY) + gt * 128, cb_row = y_row + TC_TILE_BYTES;
    const uint32_t cr_row = cb_row + TC_TILE_BYTES;  // (4:4:4 only)

    const bool issuer = __shfl_sync(0xffffffffu, (uint32_t)wg, 0) == 0;
    auto issue = [&](int tile_sel, int tab, uint32_t tmem_d, uint32_t mbar) {  // see k_transform_tc
        const int nk = REPL && tab ? 1 : 4;  // K = 16 for the chroma means
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
            const uint64_t hi = (uint64_t)0x40004040u << 32;
            const uint64_t da = hi | lds_volatile(smem_u32(&s_desc[g][tile_sel]));
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                const uint64_t db = hi | lds_volatile(smem_u32(&s_desc[T3_GROUPS][tab * 2 + s2]));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < nk) umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (s2 | k) ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
        }
    };
    // writes of this thread to the tiles become visible to the tensor core, the group meets.  Where the
    // barrier only tells the issuing warp that an accumulator has been read, the other warps arrive and go on.
    auto publish = [&](bool all_wait) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (all_wait || issuer)
            asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        else
            asm volatile("bar.arrive %0, 128;" ::"r"(1 + g) : "memory");
    };
    auto divmod = [](uint32_t n, uint32_t d, uint32_t m, uint32_t& q, uint32_t& r) {
        q = __umulhi(n, m);
        r = n - q * d;
        if (r >= d) {
            q += 1;
            r -= d;
        }
    };
    // a lane pair's MCU pair: `gm` is the index of its first (even-column) MCU in the coefficient array
    auto decode = [&](uint32_t unit_base) {
        const uint32_t lin = (unit_base + wg) * 16u + (uint32_t)(lane >> 1);
        TcUnit u;
        u.valid = unit_base + wg < a.total_units && lin < a.tc_mcus;
        const uint32_t l = u.valid ? lin : 0u;
        uint32_t f, rem, my, mx;
        divmod(l, a.tc_per_frame, a.tc_magic_frame, f, rem);
        divmod(rem, a.tc_row_len, a.tc_magic_row, my, mx);
        u.ptr = a.rgb + (size_t)f * a.frame_stride + (size_t)mx * 48;
        u.y0 = (int)my * 8;
        u.gm = f * (uint32_t)a.g.n_mcu + my * (uint32_t)a.g.mcux + 2u * mx;
        return u;
    };
    // fetch cursors: as in k_transform_tc (a pair of 8-pixel MCUs is 48 bytes = three 16-byte chunks per row)
    const uint32_t lane_share = (uint32_t)lane * (ALIGN == 16 ? 16u : 24u);
    const uint32_t ring_wr = ring + lane_share, ring_rd = ring + (uint32_t)lane * 24u;
    const int pitch_i = (int)a.pitch, img_h = a.g.H;
    const uint8_t *fc_p0 = nullptr, *fc_p1 = nullptr;
    int fc_y0 = 0, fc_y1 = 0;
    bool fc_v0 = false, fc_v1 = false;
    auto shfl_ptr = [&](const uint8_t* p, int src) {
        unsigned long long v = (unsigned long long)p;
        uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
        return (const uint8_t*)(((unsigned long long)hi << 32) | lo);
    };
    auto aim = [&](const TcUnit& u) {
        const uint8_t* row0 = u.ptr + (uint64_t)(uint32_t)u.y0 * (uint32_t)pitch_i;
        if (ALIGN == 16) {
            const int c0 = lane, c1 = 32 + (lane & 15);
            const int m0 = c0 / 3, m1 = c1 / 3;
            fc_p0 = shfl_ptr(row0, 2 * m0) + (c0 - 3 * m0) * 16;
            fc_y0 = __shfl_sync(0xffffffffu, u.y0, 2 * m0);
            fc_v0 = __shfl_sync(0xffffffffu, (int)u.valid, 2 * m0) != 0;
            fc_p1 = shfl_ptr(row0, 2 * m1) + (c1 - 3 * m1) * 16;
            fc_y1 = __shfl_sync(0xffffffffu, u.y0, 2 * m1);
            fc_v1 = __shfl_sync(0xffffffffu, (int)u.valid, 2 * m1) != 0 && lane < 16;
        } else {
            fc_p0 = row0 + half * 24;
            fc_y0 = u.y0;
            fc_v0 = u.valid;
        }
    };
    auto walk = [&](const uint8_t*& p, int& y, const uint8_t*& r0, const uint8_t*& r1) {  // even height: see k_transform_tc
        const int d = y < img_h ? pitch_i : -pitch_i;
        r0 = p;
        r1 = p + d;
        y += 2;
        p = r1 + (y == img_h ? 0 : d);
    };
    auto fetch_pair = [&](int sl) {
        const uint32_t d0 = ring_wr + (uint32_t)sl * (2 * 4 * TC_ROW_BYTES), d1 = d0 + 4 * TC_ROW_BYTES;
        const uint8_t *r0, *r1;
        walk(fc_p0, fc_y0, r0, r1);
        if (ALIGN == 16) {
            if (fc_v0) {
                cp_async<16>(d0, r0);
                cp_async<16>(d1, r1);
            }
            walk(fc_p1, fc_y1, r0, r1);
            if (fc_v1) {
                cp_async<16>(d0 + 512, r0);
                cp_async<16>(d1 + 512, r1);
            }
        } else if (fc_v0) {
#pragma unroll
            for (int j = 0; j < 24 / ALIGN; ++j) {
                cp_async<ALIGN>(d0 + j * ALIGN, r0 + j * ALIGN);
                cp_async<ALIGN>(d1 + j * ALIGN, r1 + j * ALIGN);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // Units are handed out dynamically.  A unit has only four row pairs and the copy of the next unit's
    // first rows starts at the third, before any barrier of the unit: the index of the next unit is drawn one
    // unit ahead (at the start of a unit, read after the barrier that ends its row loop).
    uint32_t base = (blockIdx.x * T3_GROUPS + g) * 4;
    if (gt == 0) s_next0[g] = stride + atomicAdd(a.unit_counter, 4u);  // (its own word: s_next is rewritten below before the
    asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");          //  other warps are known to have read this one)
    uint32_t nbase = s_next0[g];
    TcUnit cur = decode(base);
    if (base < a.total_units) {
        aim(cur);
        fetch_pair(0);
        fetch_pair(1);
    }
    while (base < a.total_units) {
        if (gt == 0) s_next[g] = stride + atomicAdd(a.unit_counter, 4u);  // the unit after next; read after the barrier below
        TcUnit nxt = cur;
        const bool valid = cur.valid;
        const uint32_t gp = valid ? cur.gm * 24u : 0xFFFFFFFFu;  // first 16-byte piece of the pair's six blocks
        const uint32_t gb0 = cur.gm * 3u;

        // ---- 8 image rows = 4 row pairs: ring -> registers -> colour conversion -> the three A tiles ----
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            if (ALIGN == 16) __syncwarp();
            const uint32_t src = ring_rd + (uint32_t)(it & 1) * (2 * 4 * TC_ROW_BYTES);
            uint32_t w0[6], w1[6];
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                uint2 v0 = lds64(src + 8 * j), v1 = lds64(src + 4 * TC_ROW_BYTES + 8 * j);
                w0[2 * j] = v0.x;
                w0[2 * j + 1] = v0.y;
                w1[2 * j] = v1.x;
                w1[2 * j + 1] = v1.y;
            }
            if (ALIGN == 16) __syncwarp();
            if (it == 2) {
                nxt = decode(nbase);
                aim(nxt);
            }
            fetch_pair(it & 1);

            Row8T o0, o1;
            csc_row8_t(w0, s_ydown, o0);
            csc_row8_t(w1, s_ydown, o1);
            const uint32_t c0 = ((uint32_t)(2 * it) ^ sw_own) << 4, c1 = ((uint32_t)(2 * it + 1) ^ sw_own) << 4;
            sts128(y_row + c0, make_uint4(luma_h2(o0.y[0], o0.y[1]), luma_h2(o0.y[2], o0.y[3]), luma_h2(o0.y[4], o0.y[5]),
                                          luma_h2(o0.y[6], o0.y[7])));
            sts128(y_row + c1, make_uint4(luma_h2(o1.y[0], o1.y[1]), luma_h2(o1.y[2], o1.y[3]), luma_h2(o1.y[4], o1.y[5]),
                                          luma_h2(o1.y[6], o1.y[7])));
            if (SUB == JB_SUB_444) {  // the top byte of a chroma T value is the sample: same packing as luma
                sts128(cb_row + c0, make_uint4(luma_h2(o0.cb[0], o0.cb[1]), luma_h2(o0.cb[2], o0.cb[3]),
                                               luma_h2(o0.cb[4], o0.cb[5]), luma_h2(o0.cb[6], o0.cb[7])));
                sts128(cb_row + c1, make_uint4(luma_h2(o1.cb[0], o1.cb[1]), luma_h2(o1.cb[2], o1.cb[3]),
                                               luma_h2(o1.cb[4], o1.cb[5]), luma_h2(o1.cb[6], o1.cb[7])));
                sts128(cr_row + c0, make_uint4(luma_h2(o0.cr[0], o0.cr[1]), luma_h2(o0.cr[2], o0.cr[3]),
                                               luma_h2(o0.cr[4], o0.cr[5]), luma_h2(o0.cr[6], o0.cr[7])));
                sts128(cr_row + c1, make_uint4(luma_h2(o1.cr[0], o1.cr[1]), luma_h2(o1.cr[2], o1.cr[3]),
                                               luma_h2(o1.cr[4], o1.cr[5]), luma_h2(o1.cr[6], o1.cr[7])));
            } else {  // performCDS (utils.cpp:113-141): the truncated mean of a 2x2 cell, written to its four samples
                uint32_t sb[4], sr[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    sb[c] = (o0.cb[2 * c] >> 24) + (o0.cb[2 * c + 1] >> 24) + (o1.cb[2 * c] >> 24) + (o1.cb[2 * c + 1] >> 24);
                    sr[c] = (o0.cr[2 * c] >> 24) + (o0.cr[2 * c + 1] >> 24) + (o1.cr[2 * c] >> 24) + (o1.cr[2 * c + 1] >> 24);
                }
                const uint32_t b01 = chroma_h2(sb[0], sb[1]), b23 = chroma_h2(sb[2], sb[3]);
                const uint32_t r01 = chroma_h2(sr[0], sr[1]), r23 = chroma_h2(sr[2], sr[3]);
                // means (it, 0..3) are K positions 4 it .. 4 it + 3: 8 bytes of the row's first (Cb) / third (Cr) 32 bytes
                const uint32_t cc = (uint32_t)(it >> 1), co = (uint32_t)(it & 1) * 8u;
                sts64(cb_row + ((cc ^ sw_own) << 4) + co, make_uint2(b01, b23));
                sts64(cb_row + (((cc + 2u) ^ sw_own) << 4) + co, make_uint2(r01, r23));
            }
        }
        publish(true);
        if (issuer) {
            issue(0, 0, tmem_d0, mbar0);  // Y
            issue(1, 1, tmem_d1, mbar1);  // Cb
        }
        const uint32_t nnbase = s_next[g];

        // read this thread's row of the accumulator, round / flag / pack, stage in the Y tile, store block `comp`
        // of both MCUs of every pair
        auto finish = [&](uint32_t tmem_d, int tab, int comp) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t t[64];
            tmem_ld64(tmem_d + lane_off, t);
            uint4* st = reinterpret_cast<uint4*>(tileY) + wg * 256;  // this warp's 32 rows of the tile
            uint32_t tl = 0, th = 0;
            if (tab == 0)
                tc_quant_stage<0>(t, a, st, lane, tl, th, 0, 0);
            else
                tc_quant_stage<1>(t, a, st, lane, tl, th, 0, 0);
            if (valid && (tl | th)) append_ties(a.tie_list, a.tie_count, a.tie_cap, gb0 + 3 * half + comp, tl, th);
            __syncwarp();
            // coalesced copy-out: staged block sb belongs to lane sb, i.e. to MCU (sb & 1) of pair sb >> 1
            const uint32_t piece0 = (uint32_t)(3 * ((lane >> 3) & 1) + comp) * 8u + (uint32_t)(lane & 7);
#pragma unroll
            for (int it8 = 0; it8 < 8; ++it8) {
                const int sb = it8 * 4 + (lane >> 3), pc = lane & 7;
                const uint32_t m_gp = __shfl_sync(0xffffffffu, gp, 4 * it8 + 2 * (lane >> 4));
                if (m_gp != 0xFFFFFFFFu) coef4[m_gp + piece0] = st[sb * 8 + (pc ^ (sb & 7))];
            }
            __syncwarp();
        };
        mbar_wait(mbar0, phase0);  // the Y tile has been consumed: it is the staging area from here on
        phase0 ^= 1;
        finish(tmem_d0, 0, 0);
        publish(false);            // every thread has read accumulator 0: it takes Cr
        if (issuer) issue(2, 1, tmem_d0, mbar0);
        mbar_wait(mbar1, phase1);
        phase1 ^= 1;
        finish(tmem_d1, 1, 1);
        mbar_wait(mbar0, phase0);
        phase0 ^= 1;
        finish(tmem_d0, 1, 2);
        cur = nxt;
        base = nbase;
        nbase = nnbase;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "n"(TC_TMEM_COLS));
}

template <int SUB, int ALIGN>
static void launch_one(const TransformArgs& a, int grid, cudaStream_t s) {
    const int smem = TW * 256 * 16 + (SUB == JB_SUB_420 ? TW * CH_WARP_WORDS * 4 : 0);
    // > 48 KB of dynamic shared memory needs the opt-in (per device, so set it on every launch: ~1 us)
    cudaFuncSetAttribute(k_transform<SUB, ALIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_transform<SUB, ALIGN><<<grid, TW * 32, smem, s>>>(a);
}

template <int SUB>
static void launch_sub(const TransformArgs& a, int align, int grid, cudaStream_t s) {
    if (align == 8)
        launch_one<SUB, 8>(a, grid, s);
    else if (align == 4)
        launch_one<SUB, 4>(a, grid, s);
    else
        launch_one<SUB, 1>(a, grid, s);
}

// MCUs the hot kernel can take: all whose 8/16-pixel columns lie inside the image, and rows
// below the image only when row mirroring reproduces the reference (no chroma cells, or an
// even height so that a mirrored row pair is again a complete 2x2 cell pair).
static int input_align(const TransformArgs& a) {
    uintptr_t bits = (uintptr_t)a.rgb | (uintptr_t)a.pitch | (uintptr_t)a.frame_stride;
    return (bits & 15) == 0 ? 16 : (bits & 7) == 0 ? 8 : (bits & 3) == 0 ? 4 : 1;
}
static bool use_tc(const TransformArgs& a) { return a.tc_mat != nullptr && input_align(a) >= 4; }

static void plan_fast(TransformArgs& a) {
    const int mcus_per_unit = a.g.sub == JB_SUB_420 ? 16 : 32;
    const bool tc = use_tc(a);
    // rows below the image are mirrored inside the hot kernel when that reproduces the reference: no chroma
    // cells (4:4:4) or an even height; the tcgen05 kernels walk row pairs and need the even height
    const bool rows_ok = (a.g.sub == JB_SUB_444 && !tc) || (a.g.H % 2 == 0);
    a.fast_mcux = a.g.W % a.g.mcu_px ? a.g.mcux - 1 : a.g.mcux;
    if (tc && a.g.sub != JB_SUB_420) a.fast_mcux &= ~1;  // the tcgen05 kernel for 8x8 MCUs takes them in pairs
    a.fast_mcuy = (a.g.H % a.g.mcu_px) && !rows_ok ? a.g.mcuy - 1 : a.g.mcuy;
    a.units_per_row = (a.fast_mcux + mcus_per_unit - 1) / mcus_per_unit;
    a.total_units = (uint32_t)a.units_per_row * (uint32_t)a.fast_mcuy * (uint32_t)a.n_frames;
}

static uint32_t div_magic32(uint32_t d) {  // floor(2^32 / d), saturated (d = 1)
    const uint64_t m = (1ull << 32) / d;
    return (uint32_t)(m > 0xFFFFFFFFull ? 0xFFFFFFFFull : m);
}

template <int SUB>
static void launch_tc3(const TransformArgs& a, int align, int grid, cudaStream_t s) {
    if (align == 16) {
        cudaFuncSetAttribute(k_transform_tc3<SUB, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, t3_smem(SUB));
        k_transform_tc3<SUB, 16><<<grid, t3_groups(SUB) * 128, t3_smem(SUB), s>>>(a);
    } else if (align == 8) {
        cudaFuncSetAttribute(k_transform_tc3<SUB, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, t3_smem(SUB));
        k_transform_tc3<SUB, 8><<<grid, t3_groups(SUB) * 128, t3_smem(SUB), s>>>(a);
    } else {
        cudaFuncSetAttribute(k_transform_tc3<SUB, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, t3_smem(SUB));
        k_transform_tc3<SUB, 4><<<grid, t3_groups(SUB) * 128, t3_smem(SUB), s>>>(a);
    }
}

// The 4:2:0 kernel with TMA-staged tiles: units are rectangles of 8 x 2 MCUs; the RGB batch is described by one 3-D
// tensor map (uint32 elements: pitch / 4 columns, H rows, n_frames frames; box = 96 x 2 x 1).  Returns false when the
// driver has no tensor-map encoder or rejects the shape (the caller then takes the cp.async kernel).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return (EncodeTiledFn)p;
    }();
    return fn;
}

void* tensor_map_encode_fn() { return (void*)tensor_map_encoder(); }

static bool launch_tma(const TransformArgs& a_in, int sms, cudaStream_t s) {
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) return false;
    TransformArgs a = a_in;
    const size_t fstride = a.n_frames > 1 ? a.frame_stride : a.pitch * (size_t)a.g.H;
    if ((a.pitch >> 2) >= (1ull << 32) || a.pitch >= (1ull << 40) || fstride >= (1ull << 40)) return false;
    CUtensorMap tm;
    const cuuint64_t dims[3] = {(cuuint64_t)(a.pitch / 4), (cuuint64_t)a.g.H, (cuuint64_t)a.n_frames};
    const cuuint64_t strides[2] = {(cuuint64_t)a.pitch, (cuuint64_t)fstride};
    const cuuint32_t box[3] = {96, 2, 1}, estr[3] = {1, 1, 1};
    if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(a.rgb), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    const uint32_t units_x = (uint32_t)((a.fast_mcux + 7) / 8), units_y = (uint32_t)((a.fast_mcuy + 1) / 2);
    a.tc_row_len = units_x;
    a.tc_per_frame = units_x * units_y;
    a.total_units = a.tc_per_frame * (uint32_t)a.n_frames;
    if (!a.total_units) return false;
    a.tc_magic_frame = div_magic32(a.tc_per_frame);
    a.tc_magic_row = div_magic32(a.tc_row_len);
    const int needg = (int)((a.total_units + 4 * TC_GROUPS - 1) / (4 * TC_GROUPS));
    cudaFuncSetAttribute(k_transform_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
    k_transform_tma<<<needg < sms ? needg : sms, TC_GROUPS * 128, TC_SMEM, s>>>(a, tm);
    return true;
}

// NV12-style input: the MCUs that lie inside the image through the tcgen05 kernel (16-byte aligned planes, pitches and
// frame strides; JB_FLAG_FMA_DCT off), the rest -- and everything otherwise -- through the CUDA-core kernel.
static bool nv12_plan_fast(TransformArgs& a) {
    a.fast_mcux = a.fast_mcuy = 0;
    const uintptr_t bits = (uintptr_t)a.rgb | (uintptr_t)a.uv | a.pitch | a.pitch_uv | (a.n_frames > 1 ? a.frame_stride | a.frame_stride_uv : 0);
    if (a.tc_mat == nullptr || (bits & 15u) != 0 || a.g.W < 16 || a.g.H < 16) return false;
    a.fast_mcux = a.g.W / 16;   // complete MCU columns; rows below the image are mirrored by the kernel itself
    a.fast_mcuy = a.g.mcuy;
    return true;
}

int launch_transform_nv12(const TransformArgs& a_in, cudaStream_t s) {
    TransformArgs a = a_in;
    const size_t total = (size_t)a.g.n_mcu * 6 * (size_t)a.n_frames;
    if (!total) return 0;
    if (nv12_plan_fast(a)) {
        int sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        a.tc_row_len = (uint32_t)a.fast_mcux;
        a.tc_per_frame = a.tc_row_len * (uint32_t)a.fast_mcuy;
        a.tc_mcus = a.tc_per_frame * (uint32_t)a.n_frames;
        a.total_units = (a.tc_mcus + 15) / 16;
        a.tc_magic_frame = div_magic32(a.tc_per_frame);
        a.tc_magic_row = div_magic32(a.tc_row_len);
        const int needg = (int)((a.total_units + 4 * TC_GROUPS - 1) / (4 * TC_GROUPS));
        cudaFuncSetAttribute(k_transform_tc_nv12, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
        k_transform_tc_nv12<<<needg < sms ? needg : sms, TC_GROUPS * 128, TC_SMEM, s>>>(a);
        return 1;
    }
    const size_t g = (total + 127) / 128;
    k_transform_nv12<<<(int)(g > 148 * 64 ? 148 * 64 : g), 128, 0, s>>>(a);
    return 1;
}

// the MCUs the tcgen05 kernel left out (a partial last MCU row / column)
int launch_transform_nv12_edge(const TransformArgs& a_in, cudaStream_t s) {
    TransformArgs a = a_in;
    const size_t total = (size_t)a.g.n_mcu * 6 * (size_t)a.n_frames;
    if (!total || !nv12_plan_fast(a) || (size_t)a.fast_mcux * a.fast_mcuy == (size_t)a.g.n_mcu) return 0;
    const size_t g = (total + 127) / 128;
    k_transform_nv12<<<(int)(g > 148 * 64 ? 148 * 64 : g), 128, 0, s>>>(a);
    return 1;
}

int launch_transform(const TransformArgs& a_in, cudaStream_t s) {
    TransformArgs a = a_in;
    plan_fast(a);
    if (!a.total_units) return 0;
    const int align16 = input_align(a), align = align16 == 16 ? 8 : align16;
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);  // the context's device (jb_api sets it before every launch), not device 0
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (use_tc(a)) {  // tcgen05 kernels: units are runs of 16 MCUs (4:2:0) or 16 MCU pairs (8x8 MCUs)
        const bool is420 = a.g.sub == JB_SUB_420;
        a.tc_row_len = (uint32_t)(is420 ? a.fast_mcux : a.fast_mcux / 2);
        a.tc_per_frame = a.tc_row_len * (uint32_t)a.fast_mcuy;
        a.tc_mcus = a.tc_per_frame * (uint32_t)a.n_frames;
        a.total_units = (a.tc_mcus + 15) / 16;
        if (!a.total_units) return 0;
        a.tc_magic_frame = div_magic32(a.tc_per_frame);
        a.tc_magic_row = div_magic32(a.tc_row_len);
        const int groups = is420 ? TC_GROUPS : t3_groups(a.g.sub);
        int needg = (int)((a.total_units + 4 * groups - 1) / (4 * groups));
        int gridg = needg < sms ? needg : sms;
        if (!is420) {
            if (a.g.sub == JB_SUB_444)
                launch_tc3<JB_SUB_444>(a, align16, gridg, s);
            else
                launch_tc3<JB_SUB_REPL420>(a, align16, gridg, s);
        } else if (align16 == 16 && a.use_tma && launch_tma(a, sms, s)) {
            // TMA-staged tiles (k_transform_tma): 16-byte aligned base / pitch / frame stride
        } else if (align16 == 16) {
            cudaFuncSetAttribute(k_transform_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
            k_transform_tc<16><<<gridg, TC_GROUPS * 128, TC_SMEM, s>>>(a);
        } else if (align == 8) {
            cudaFuncSetAttribute(k_transform_tc<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
            k_transform_tc<8><<<gridg, TC_GROUPS * 128, TC_SMEM, s>>>(a);
        } else {
            cudaFuncSetAttribute(k_transform_tc<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM);
            k_transform_tc<4><<<gridg, TC_GROUPS * 128, TC_SMEM, s>>>(a);
        }
        return 1;
    }
    int need = (int)((a.total_units + TW - 1) / TW);
    int grid = need < sms * (16 / TW) ? need : sms * (16 / TW);
    if (a.g.sub == JB_SUB_420)
        launch_sub<JB_SUB_420>(a, align, grid, s);
    else if (a.g.sub == JB_SUB_REPL420)
        launch_sub<JB_SUB_REPL420>(a, align, grid, s);
    else
        launch_sub<JB_SUB_444>(a, align, grid, s);
    return 1;
}

int launch_transform_edge(const TransformArgs& a_in, cudaStream_t s) {
    TransformArgs a = a_in;
    plan_fast(a);
    size_t edge = ((size_t)(a.g.mcux - a.fast_mcux) * (size_t)a.g.mcuy + (a.fast_mcuy < a.g.mcuy ? (size_t)a.fast_mcux : 0)) *
                  (size_t)a.g.bpm * (size_t)a.n_frames;
    if (!edge) return 0;
    size_t g = (edge + 63) / 64;
    k_transform_edge<<<(int)(g > 148 * 32 ? 148 * 32 : g), 64, 0, s>>>(a);
    return 1;
}

// ---------------------------------------------------------------- fix-up ----
// One warp per batch of FIX_BATCH listed coefficients.  For each entry in turn the lanes produce the 64
// products sample x cos x cos in parallel (two each) into shared memory; then lane L adds the 64
// terms of entry L in the reference's order (utils.cpp:314-347 with the block read from a copy,
// then utils.cpp:454-467), in binary64 with unfused multiplies and adds: the serial chains of 32
// entries run side by side instead of one chain occupying a whole warp.
#ifndef FIX_BATCH
#define FIX_BATCH 8
#endif
constexpr int FIX_WARPS = 4;  // the kernel is latency-bound: small batches keep 64 warps per SM resident
__global__ void __launch_bounds__(FIX_WARPS * 32) k_fixup(const __grid_constant__ FixupArgs a) {
    __shared__ double s_term[FIX_WARPS][FIX_BATCH][65];  // [entry of the batch][term], padded: conflict-free both ways
    __shared__ double s_smp[FIX_WARPS][64];              // level-shifted samples of the block being replayed
    uint32_t n = *a.tie_count;
    if (n > a.tie_cap) n = a.tie_cap;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t warps = gridDim.x * FIX_WARPS;
    const uint32_t bpf = (uint32_t)a.g.n_mcu * (uint32_t)a.g.bpm;
    for (uint32_t e0 = (blockIdx.x * FIX_WARPS + w) * FIX_BATCH; e0 < n; e0 += warps * FIX_BATCH) {
        const int n_here = (int)min((uint32_t)FIX_BATCH, n - e0);
        const uint32_t my_entry = lane < n_here ? a.tie_list[e0 + lane] : 0u;
        // lane L decodes entry L once (block -> frame, MCU, component, first sample, zigzag position); the
        // fields travel to the whole warp by shuffles when the entry's turn comes
        uint32_t d_x0, d_y0, d_misc, d_f;  // (two words for the origin: dimensions go up to 2^24 under JB_FLAG_CLAMP_SOF)
        int my_comp, my_nat;
        {
            const uint32_t gblock = my_entry >> 6, k = my_entry & 63;
            // (host magics m = ceil(2^52 / d): exact for x, d < 2^26, as in the entropy coder)
            const uint32_t f = (uint32_t)__umul64hi((uint64_t)gblock << 12, a.m_bpf), rb = gblock - f * bpf;
            uint32_t mcu, blk;
            if (a.g.bpm == 6) {
                mcu = rb / 6u;
                blk = rb - mcu * 6u;
            } else {
                mcu = rb / 3u;
                blk = rb - mcu * 3u;
            }
            const uint32_t my = (uint32_t)__umul64hi((uint64_t)mcu << 12, a.m_mcux), mx = mcu - my * (uint32_t)a.g.mcux;
            uint32_t x0, y0, step;
            if (a.g.sub == JB_SUB_420) {
                my_comp = blk < 4 ? 0 : (int)blk - 3;
                x0 = mx * 16 + (blk < 4 ? (blk & 1) * 8 : 0);
                y0 = my * 16 + (blk < 4 ? (blk >> 1) * 8 : 0);
                step = blk < 4 ? 1 : 2;
            } else {
                my_comp = (int)blk;
                x0 = mx * 8;
                y0 = my * 8;
                step = 1;
            }
            my_nat = c_zz[k];
            d_x0 = x0;
            d_y0 = y0;
            // bit 10: full-resolution block (luma, or any component of 4:4:4) without mirrored columns, RGB input: the
            // row-per-lane path below
            const uint32_t fast = (my_comp == 0 || a.g.sub == JB_SUB_444) && !a.uv && x0 + 8 <= (uint32_t)a.g.W ? 0x400u : 0u;
            d_misc = (uint32_t)my_comp | (step << 2) | ((uint32_t)my_nat << 4) | fast;
            d_f = f;
        }
        // ---- products.  Full-resolution blocks that need no mirrored column (luma; every component of 4:4:4): a lane per block ROW, eight
        // lanes per entry, four entries at a time -- one address, 24 contiguous bytes and eight luma values per lane
        // instead of two pixels with an address (and mirror tests) each.  The others (subsampled chroma: 2x2 cells;
        // blocks on the right edge; NV12 planes) take the pixel-by-pixel path below, an entry at a time.
#pragma unroll 1
        for (int g4 = 0; g4 < FIX_BATCH / 4; ++g4) {
            const int j = 4 * g4 + (lane >> 3), row = lane & 7;
            const uint32_t misc = __shfl_sync(0xffffffffu, d_misc, j);
            const size_t fr = (size_t)__shfl_sync(0xffffffffu, d_f, j);
            const int x0 = (int)__shfl_sync(0xffffffffu, d_x0, j), y0 = (int)__shfl_sync(0xffffffffu, d_y0, j);
            const bool fast = j < n_here && (misc & 0x400u) != 0;
            if (fast) {
                const int nat = (int)((misc >> 4) & 63u), v = nat >> 3, u = nat & 7, comp = (int)(misc & 3u);
                const uint8_t* p = a.rgb + fr * a.frame_stride + (size_t)mirror(y0 + row, a.g.H) * a.pitch + (size_t)x0 * 3;
                uint32_t px[6];
                if (a.rgb_align4)
                    load24<4>(p, px);
                else
                    load24<1>(p, px);
                const double cv = a.costab[v * 8 + row];
#pragma unroll
                for (int x = 0; x < 8; ++x) {
                    const uint32_t cr_ = byte24(px, 3 * x), cg_ = byte24(px, 3 * x + 1), cb_ = byte24(px, 3 * x + 2);
                    const uint32_t y = comp == 0 ? csc_y(cr_, cg_, cb_, a.ydown) : comp == 1 ? csc_cb(cr_, cg_, cb_) : csc_cr(cr_, cg_, cb_);
                    const double smp = __dsub_rn((double)y, 128.0);                                        // utils.cpp:190
                    s_term[w][j][row * 8 + x] = a.inplace_dct ? smp : __dmul_rn(__dmul_rn(smp, a.costab[u * 8 + x]), cv);  // utils.cpp:330
                }
            }
        }
        uint32_t have_block = 0xFFFFFFFFu;  // consecutive entries of one block (append_ties) share its samples
        for (int j = 0; j < n_here; ++j) {
            const uint32_t misc = __shfl_sync(0xffffffffu, d_misc, j), gblock = __shfl_sync(0xffffffffu, my_entry, j) >> 6;
            if (misc & 0x400u) continue;  // (warp-uniform) done above
            const size_t fr = (size_t)__shfl_sync(0xffffffffu, d_f, j);
            const int x0 = (int)__shfl_sync(0xffffffffu, d_x0, j), y0 = (int)__shfl_sync(0xffffffffu, d_y0, j);
            const int comp = (int)(misc & 3u), step = (int)((misc >> 2) & 3u);
            const int nat = (int)((misc >> 4) & 63u), v = nat >> 3, u = nat & 7;
            if (gblock != have_block) {  // (warp-uniform)
                Image im{a.rgb + fr * a.frame_stride, a.pitch, a.g.W, a.g.H, a.ydown};
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = lane + 32 * h, x = i & 7, y = i >> 3;
                    const double smp = a.uv ? (double)nv12_sample(im.base, a.pitch, a.uv + fr * a.frame_stride_uv, a.pitch_uv, a.g.W, a.g.H,
                                                                   x0 + x * step, y0 + y * step, comp)
                                            : (double)sample_at(im, x0 + x * step, y0 + y * step, comp, a.g.sub != JB_SUB_444);  // utils.cpp:236
                    s_smp[w][i] = __dsub_rn(smp, 128.0);                                                                  // utils.cpp:190
                }
                __syncwarp();
                have_block = gblock;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = lane + 32 * h, x = i & 7, y = i >> 3;
                const double smp = s_smp[w][i];
                s_term[w][j][i] = a.inplace_dct ? smp : __dmul_rn(__dmul_rn(smp, a.costab[u * 8 + x]), a.costab[v * 8 + y]);  // utils.cpp:330
            }
        }
        __syncwarp();
        if (lane < n_here) {
            const int v = my_nat >> 3, u = my_nat & 7;
            double sum = 0.0;  // y outer, x inner = index order 0..63
            if (!a.inplace_dct) {
#pragma unroll 8
                for (int i = 0; i < 64; ++i) sum = __dadd_rn(sum, s_term[w][lane][i]);
                sum = __dmul_rn(sum, a.scale[u * 8 + v]);                                                 // utils.cpp:336
            } else {
                // Q1 (utils.cpp:314-347): the block is overwritten while it is read.  s_term holds the level-shifted
                // samples here; run the reference's loop (outputs u outer / v inner, each stored at row v, column u
                // of the block it goes on reading) up to the flagged output.
                double* blk = s_term[w][lane];
                for (int uu = 0; uu <= u; ++uu)
                    for (int vv = 0; vv < (uu == u ? v + 1 : 8); ++vv) {
                        double acc = 0.0;
                        for (int i = 0; i < 64; ++i)
                            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(blk[i], a.costab[uu * 8 + (i & 7)]), a.costab[vv * 8 + (i >> 3)]));
                        acc = __dmul_rn(acc, a.scale[uu * 8 + vv]);
                        blk[vv * 8 + uu] = acc;
                        sum = acc;
                    }
            }
            double q = (double)a.qt.q[my_comp ? 1 : 0][my_nat];
            double r = round(__ddiv_rn(sum, q));                                                          // utils.cpp:460
            a.coef[(size_t)my_entry] = (int16_t)(int)r;  // entry = block * 64 + zigzag position            utils.cpp:490
        }
        __syncwarp();
    }
}

int launch_fixup(const FixupArgs& a_in, cudaStream_t s) {
    FixupArgs a = a_in;
    a.m_bpf = ((1ull << 52) + (uint64_t)a.g.n_mcu * a.g.bpm - 1) / ((uint64_t)a.g.n_mcu * a.g.bpm);
    a.m_mcux = ((1ull << 52) + (uint64_t)a.g.mcux - 1) / (uint64_t)a.g.mcux;
    a.rgb_align4 = ((reinterpret_cast<uintptr_t>(a.rgb) | a.pitch | a.frame_stride) & 3u) == 0 ? 1 : 0;
    k_fixup<<<148 * 16, FIX_WARPS * 32, 0, s>>>(a);
    return 1;
}

}  // namespace jb
