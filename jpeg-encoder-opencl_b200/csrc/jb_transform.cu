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
// both DCT passes, quantisation and the zigzag permutation need no shuffles and
// no shared memory.  Results are staged through a per-warp, XOR-swizzled 4 KB
// shared-memory tile so that the HBM stores are coalesced 128-bit stores of
// whole 128-byte blocks.  Algorithmic HBM traffic: 3 B/px read + 2 B/sample
// written = 6 B/px (4:2:0) or 9 B/px (4:4:4, replicated 4:2:0).
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
#pragma unroll
    for (int p = 0; p < 8; ++p)
        st[lane * 8 + (p ^ (lane & 7))] = make_uint4(wd[4 * p], wd[4 * p + 1], wd[4 * p + 2], wd[4 * p + 3]);
}

__device__ __forceinline__ void append_ties(const TransformArgs& a, uint32_t gblock, uint32_t lo, uint32_t hi) {
    while (lo | hi) {
        int k;
        if (lo) {
            k = __ffs(lo) - 1;
            lo &= lo - 1;
        } else {
            k = 31 + __ffs(hi);
            hi &= hi - 1;
        }
        uint32_t idx = atomicAdd(a.tie_count, 1u);
        if (idx < a.tie_cap) a.tie_list[idx] = gblock * 64u + (uint32_t)k;
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

__device__ __forceinline__ uint32_t pack_s8(uint32_t acc, int val, int pos) {
    return acc | (((uint32_t)val & 0xFFu) << (8 * pos));
}
__device__ __forceinline__ float unpack_s8(uint32_t w, int pos) { return (float)(int)(int8_t)(w >> (8 * pos)); }

// ------------------------------------------------------------------ 4:2:0 --
// Lane = one 8-pixel-wide, 16-row half of a 16x16 MCU: two luma blocks, then
// (after exchanging chroma halves with the partner lane) one chroma block.
template <int ALIGN>
__device__ __forceinline__ void load_half_420(const Image& im, int px, int py, int h, bool interior, float (&v)[64],
                                              uint32_t (&cbq)[8], uint32_t (&crq)[8]) {
    uint32_t pcb[4], pcr[4];
    if (interior) {
        const uint8_t* row = im.base + (size_t)(py + h * 8) * im.pitch + (size_t)px * 3;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint32_t w[6];
            load24<ALIGN>(row + (size_t)r * im.pitch, w);
            uint32_t hcb[4], hcr[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t R = byte24(w, 3 * i), G = byte24(w, 3 * i + 1), B = byte24(w, 3 * i + 2);
                v[r * 8 + i] = (float)((int)csc_y(R, G, B, im.ydown) - 128);
                uint32_t cb = csc_cb(R, G, B), cr = csc_cr(R, G, B);
                if (i & 1) {
                    hcb[i >> 1] += cb;
                    hcr[i >> 1] += cr;
                } else {
                    hcb[i >> 1] = cb;
                    hcr[i >> 1] = cr;
                }
            }
            if (r & 1) {
                uint32_t qb = 0, qr = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    qb = pack_s8(qb, (int)((pcb[c] + hcb[c]) >> 2) - 128, c);
                    qr = pack_s8(qr, (int)((pcr[c] + hcr[c]) >> 2) - 128, c);
                }
                cbq[h * 4 + (r >> 1)] = qb;
                crq[h * 4 + (r >> 1)] = qr;
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    pcb[c] = hcb[c];
                    pcr[c] = hcr[c];
                }
            }
        }
    } else {
        // fully unrolled so that v[] keeps static indices (stays in registers)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint32_t qb = 0, qr = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t Y, Cb, Cr;
                ycc_at<true>(im, px + i, py + h * 8 + r, Y, Cb, Cr);
                v[r * 8 + i] = (float)((int)Y - 128);
                // chroma sample (ci, cj) of the 4:2:0 planes = replicated plane at (2ci, 2cj)
                if ((r & 1) == 0 && (i & 1) == 0) {
                    qb = pack_s8(qb, (int)Cb - 128, i >> 1);
                    qr = pack_s8(qr, (int)Cr - 128, i >> 1);
                }
            }
            if ((r & 1) == 0) {
                cbq[h * 4 + (r >> 1)] = qb;
                crq[h * 4 + (r >> 1)] = qr;
            }
        }
    }
}

template <int ALIGN>
__device__ __forceinline__ void unit_420(const TransformArgs& a, const Image& im, size_t mcu_g0, int mcu_x0, int my,
                                         uint4* st, int lane) {
    const int m = lane >> 1, half = lane & 1;
    const int mcus_valid = min(16, a.g.mcux - mcu_x0);
    const bool valid = m < mcus_valid;
    const int px = (mcu_x0 + m) * 16 + half * 8, py = my * 16;
    const bool interior = px + 8 <= a.g.W && py + 16 <= a.g.H;
    const uint32_t gb0 = (uint32_t)(mcu_g0 + m) * 6u;
    uint4* coef4 = reinterpret_cast<uint4*>(a.coef);
    float v[64];
    uint32_t cbq[8], crq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cbq[j] = crq[j] = 0;

#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (valid) {
            load_half_420<ALIGN>(im, px, py, h, interior, v, cbq, crq);
            fdct2d(v);
            uint32_t tl = 0, th = 0;
            quant_stage<0>(v, a, st, lane, tl, th);
            if (tl | th) append_ties(a, gb0 + 2 * h + half, tl, th);
        }
        __syncwarp();
        copy_out<6>(st, coef4, mcu_g0, mcus_valid, 2 * h, lane);
        __syncwarp();
    }
    // exchange chroma halves: the left lane builds Cb, the right lane builds Cr
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t send = half ? cbq[j] : crq[j];
        uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
        uint32_t lo = half ? recv : cbq[j];
        uint32_t hi = half ? crq[j] : recv;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[j * 8 + i] = unpack_s8(lo, i);
            v[j * 8 + 4 + i] = unpack_s8(hi, i);
        }
    }
    if (valid) {
        fdct2d(v);
        uint32_t tl = 0, th = 0;
        quant_stage<1>(v, a, st, lane, tl, th);
        if (tl | th) append_ties(a, gb0 + 4 + half, tl, th);
    }
    __syncwarp();
    copy_out<6>(st, coef4, mcu_g0, mcus_valid, 4, lane);
    __syncwarp();
}

// ------------------------------------------------- 4:4:4 / replicated 4:2:0 --
// Lane = one 8x8 MCU: Y, Cb, Cr blocks in turn; chroma is parked as packed bytes.
template <int ALIGN, bool CDS>
__device__ __forceinline__ void load_mcu_444(const Image& im, int px, int py, bool interior, float (&v)[64],
                                             uint32_t (&cq)[2][8][2]) {
    if (interior) {
        const uint8_t* row = im.base + (size_t)py * im.pitch + (size_t)px * 3;
        uint32_t pcb[4], pcr[4];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint32_t w[6];
            load24<ALIGN>(row + (size_t)r * im.pitch, w);
            uint32_t cb[8], cr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t R = byte24(w, 3 * i), G = byte24(w, 3 * i + 1), B = byte24(w, 3 * i + 2);
                v[r * 8 + i] = (float)((int)csc_y(R, G, B, im.ydown) - 128);
                cb[i] = csc_cb(R, G, B);
                cr[i] = csc_cr(R, G, B);
            }
            if (!CDS) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if ((i & 3) == 0) cq[0][r][i >> 2] = cq[1][r][i >> 2] = 0;
                    cq[0][r][i >> 2] = pack_s8(cq[0][r][i >> 2], (int)cb[i] - 128, i & 3);
                    cq[1][r][i >> 2] = pack_s8(cq[1][r][i >> 2], (int)cr[i] - 128, i & 3);
                }
            } else if ((r & 1) == 0) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    pcb[c] = cb[2 * c] + cb[2 * c + 1];
                    pcr[c] = cr[2 * c] + cr[2 * c + 1];
                }
            } else {
                uint32_t qb[2] = {0, 0}, qr[2] = {0, 0};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    int mb = (int)((pcb[c] + cb[2 * c] + cb[2 * c + 1]) >> 2) - 128;
                    int mr = (int)((pcr[c] + cr[2 * c] + cr[2 * c + 1]) >> 2) - 128;
                    qb[c >> 1] = pack_s8(pack_s8(qb[c >> 1], mb, (2 * c) & 3), mb, (2 * c + 1) & 3);
                    qr[c >> 1] = pack_s8(pack_s8(qr[c >> 1], mr, (2 * c) & 3), mr, (2 * c + 1) & 3);
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    cq[0][r - 1][k] = cq[0][r][k] = qb[k];
                    cq[1][r - 1][k] = cq[1][r][k] = qr[k];
                }
            }
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint32_t qb[2] = {0, 0}, qr[2] = {0, 0};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t Y, Cb, Cr;
                ycc_at<CDS>(im, px + i, py + r, Y, Cb, Cr);
                v[r * 8 + i] = (float)((int)Y - 128);
                qb[i >> 2] = pack_s8(qb[i >> 2], (int)Cb - 128, i & 3);
                qr[i >> 2] = pack_s8(qr[i >> 2], (int)Cr - 128, i & 3);
            }
            cq[0][r][0] = qb[0];
            cq[0][r][1] = qb[1];
            cq[1][r][0] = qr[0];
            cq[1][r][1] = qr[1];
        }
    }
}

template <int ALIGN, bool CDS>
__device__ __forceinline__ void unit_444(const TransformArgs& a, const Image& im, size_t mcu_g0, int mcu_x0, int my,
                                         uint4* st, int lane) {
    const int mcus_valid = min(32, a.g.mcux - mcu_x0);
    const bool valid = lane < mcus_valid;
    const int px = (mcu_x0 + lane) * 8, py = my * 8;
    const bool interior = px + 8 <= a.g.W && py + 8 <= a.g.H;
    const uint32_t gb0 = (uint32_t)(mcu_g0 + lane) * 3u;
    uint4* coef4 = reinterpret_cast<uint4*>(a.coef);
    float v[64];
    uint32_t cq[2][8][2];
    if (valid) {
        load_mcu_444<ALIGN, CDS>(im, px, py, interior, v, cq);
        fdct2d(v);
        uint32_t tl = 0, th = 0;
        quant_stage<0>(v, a, st, lane, tl, th);
        if (tl | th) append_ties(a, gb0, tl, th);
    }
    __syncwarp();
    copy_out<3>(st, coef4, mcu_g0, mcus_valid, 0, lane);
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (valid) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[r * 8 + i] = unpack_s8(cq[c][r][i >> 2], i & 3);
            fdct2d(v);
            uint32_t tl = 0, th = 0;
            quant_stage<1>(v, a, st, lane, tl, th);
            if (tl | th) append_ties(a, gb0 + 1 + c, tl, th);
        }
        __syncwarp();
        copy_out<3>(st, coef4, mcu_g0, mcus_valid, 1 + c, lane);
        __syncwarp();
    }
}

template <int SUB, int ALIGN>
__global__ void __launch_bounds__(128, 4) k_transform(const __grid_constant__ TransformArgs a) {
    __shared__ uint4 stage[4][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* st = stage[warp];
    const uint32_t units_per_frame = (uint32_t)a.units_per_row * (uint32_t)a.g.mcuy;
    const int mcus_per_unit = SUB == JB_SUB_420 ? 16 : 32;
    for (uint32_t unit = blockIdx.x * 4 + warp; unit < a.total_units; unit += gridDim.x * 4) {
        uint32_t f = unit / units_per_frame, rem = unit - f * units_per_frame;
        int my = (int)(rem / (uint32_t)a.units_per_row), ux = (int)(rem - (uint32_t)my * (uint32_t)a.units_per_row);
        Image im{a.rgb + (size_t)f * a.frame_stride, a.pitch, a.g.W, a.g.H, a.ydown};
        int mcu_x0 = ux * mcus_per_unit;
        size_t mcu_g0 = (size_t)f * (size_t)a.g.n_mcu + (size_t)my * (size_t)a.g.mcux + (size_t)mcu_x0;
        if (SUB == JB_SUB_420)
            unit_420<ALIGN>(a, im, mcu_g0, mcu_x0, my, st, lane);
        else
            unit_444<ALIGN, SUB == JB_SUB_REPL420>(a, im, mcu_g0, mcu_x0, my, st, lane);
    }
}

template <int SUB>
static void launch_sub(const TransformArgs& a, int align, int grid, cudaStream_t s) {
    if (align == 8)
        k_transform<SUB, 8><<<grid, 128, 0, s>>>(a);
    else if (align == 4)
        k_transform<SUB, 4><<<grid, 128, 0, s>>>(a);
    else
        k_transform<SUB, 1><<<grid, 128, 0, s>>>(a);
}

int launch_transform(const TransformArgs& a_in, cudaStream_t s) {
    TransformArgs a = a_in;
    const int mcus_per_unit = a.g.sub == JB_SUB_420 ? 16 : 32;
    a.units_per_row = (a.g.mcux + mcus_per_unit - 1) / mcus_per_unit;
    a.total_units = (uint32_t)a.units_per_row * (uint32_t)a.g.mcuy * (uint32_t)a.n_frames;
    if (a.total_units == 0) return 0;
    uintptr_t bits = (uintptr_t)a.rgb | (uintptr_t)a.pitch | (uintptr_t)a.frame_stride;
    int align = (bits & 7) == 0 ? 8 : (bits & 3) == 0 ? 4 : 1;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int need = (int)((a.total_units + 3) / 4);
    int grid = need < sms * 16 ? need : sms * 16;
    if (a.g.sub == JB_SUB_420)
        launch_sub<JB_SUB_420>(a, align, grid, s);
    else if (a.g.sub == JB_SUB_REPL420)
        launch_sub<JB_SUB_REPL420>(a, align, grid, s);
    else
        launch_sub<JB_SUB_444>(a, align, grid, s);
    return 1;
}

// ---------------------------------------------------------------- fix-up ----
// One thread per listed coefficient: the reference's formula in the reference's
// operation order (utils.cpp:314-347 with the block read from a copy, then
// utils.cpp:454-467), in binary64 with unfused multiplies and adds.
__global__ void k_fixup(const __grid_constant__ FixupArgs a) {
    uint32_t n = *a.tie_count;
    if (n > a.tie_cap) n = a.tie_cap;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        uint32_t entry = a.tie_list[e];
        uint32_t gblock = entry >> 6, k = entry & 63;
        uint32_t bpf = (uint32_t)a.g.n_mcu * (uint32_t)a.g.bpm;
        uint32_t f = gblock / bpf, rb = gblock - f * bpf;
        uint32_t mcu = rb / (uint32_t)a.g.bpm, blk = rb - mcu * (uint32_t)a.g.bpm;
        int my = (int)(mcu / (uint32_t)a.g.mcux), mx = (int)(mcu - (uint32_t)my * (uint32_t)a.g.mcux);
        int comp, x0, y0, step;
        if (a.g.sub == JB_SUB_420) {
            if (blk < 4) {
                comp = 0;
                x0 = mx * 16 + (int)(blk & 1) * 8;
                y0 = my * 16 + (int)(blk >> 1) * 8;
                step = 1;
            } else {
                comp = (int)blk - 3;
                x0 = mx * 16;
                y0 = my * 16;
                step = 2;
            }
        } else {
            comp = (int)blk;
            x0 = mx * 8;
            y0 = my * 8;
            step = 1;
        }
        Image im{a.rgb + (size_t)f * a.frame_stride, a.pitch, a.g.W, a.g.H, a.ydown};
        int nat = c_zz[k], v = nat >> 3, u = nat & 7;
        double sum = 0.0;
        for (int y = 0; y < 8; ++y)
            for (int x = 0; x < 8; ++x) {
                uint32_t Y, Cb, Cr;
                if (a.g.sub == JB_SUB_444)
                    ycc_at<false>(im, x0 + x * step, y0 + y * step, Y, Cb, Cr);
                else
                    ycc_at<true>(im, x0 + x * step, y0 + y * step, Y, Cb, Cr);
                double smp = (double)(comp == 0 ? Y : comp == 1 ? Cb : Cr);  // utils.cpp:236
                smp = __dsub_rn(smp, 128.0);                                   // utils.cpp:190
                double t = __dmul_rn(__dmul_rn(smp, a.costab[u * 8 + x]), a.costab[v * 8 + y]);
                sum = __dadd_rn(sum, t);                                       // utils.cpp:330
            }
        sum = __dmul_rn(sum, a.scale[u * 8 + v]);                              // utils.cpp:336
        double q = (double)a.qt.q[comp ? 1 : 0][nat];
        double r = round(__ddiv_rn(sum, q));                                   // utils.cpp:460
        a.coef[(size_t)gblock * 64 + k] = (int16_t)(int)r;                     // utils.cpp:490
    }
}

int launch_fixup(const FixupArgs& a, cudaStream_t s) {
    k_fixup<<<296, 128, 0, s>>>(a);
    return 1;
}

}  // namespace jb
