// Device helpers: the reference's pixel pipeline (CSC -> chroma 2x2 mean on the
// unpadded image -> mirror padding) evaluated at an arbitrary padded coordinate.
// Used by the edge path of the fused kernel, the binary64 fix-up kernel and
// the staged kernels.  Reference: src/utils.cpp:92-141, 199-233 in the order of
// src/OpenCLProject_JpegEncoder.cpp:59-120.
#pragma once
#include "jb_internal.h"

namespace jb {

__host__ __device__ constexpr int zz_nat(int k) {
    constexpr unsigned char t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return t[k];
}

// zigzag position -> natural index, for run-time indices
static __device__ __constant__ unsigned char c_zz[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Image {
    const uint8_t* base;  // frame base (RGB8 AoS)
    size_t pitch;
    int W, H;
    const uint32_t* ydown;
};

// addReversedPadding (utils.cpp:211-233): padded x >= W reads W - (x - W + 1).
__device__ __forceinline__ int mirror(int x, int n) { return x < n ? x : 2 * n - x - 1; }

__device__ __forceinline__ void rgb_at(const Image& im, int sx, int sy, uint32_t& r, uint32_t& g, uint32_t& b) {
    const uint8_t* p = im.base + (size_t)sy * im.pitch + (size_t)sx * 3;
    r = __ldg(p);
    g = __ldg(p + 1);
    b = __ldg(p + 2);
}

// Y, Cb, Cr (bytes, as held in the reference's padded ppm_t) at padded (x, y).
// CDS: apply performCDS (utils.cpp:113-141): a 2x2 cell whose four pixels are
// inside the unpadded image carries the truncated mean; an odd last row or
// column keeps its own chroma.
template <bool CDS>
__device__ __noinline__ void ycc_at(const Image& im, int x, int y, uint32_t& Y, uint32_t& Cb, uint32_t& Cr) {
    int sx = mirror(x, im.W), sy = mirror(y, im.H);
    uint32_t r, g, b;
    rgb_at(im, sx, sy, r, g, b);
    Y = csc_y(r, g, b, im.ydown);
    int cx = sx & ~1, cy = sy & ~1;
    if (CDS && cx + 1 < im.W && cy + 1 < im.H) {
        uint32_t scb = 0, scr = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                rgb_at(im, cx + i, cy + j, r, g, b);
                scb += csc_cb(r, g, b);
                scr += csc_cr(r, g, b);
            }
        Cb = scb >> 2;
        Cr = scr >> 2;
    } else {
        Cb = csc_cb(r, g, b);
        Cr = csc_cr(r, g, b);
    }
}

// One component (0 Y, 1 Cb, 2 Cr) of the same pipeline: what the binary64 replay needs.  `comp` and `cds`
// are uniform over the warp there, so only the loads and the arithmetic of that component are issued.
__device__ __forceinline__ uint32_t sample_at(const Image& im, int x, int y, int comp, bool cds) {
    const int sx = mirror(x, im.W), sy = mirror(y, im.H);
    uint32_t r, g, b;
    if (comp == 0) {
        rgb_at(im, sx, sy, r, g, b);
        return csc_y(r, g, b, im.ydown);
    }
    const int cx = sx & ~1, cy = sy & ~1;
    if (cds && cx + 1 < im.W && cy + 1 < im.H) {
        uint32_t s = 0;
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                rgb_at(im, cx + i, cy + j, r, g, b);
                s += comp == 1 ? csc_cb(r, g, b) : csc_cr(r, g, b);
            }
        return s >> 2;
    }
    rgb_at(im, sx, sy, r, g, b);
    return comp == 1 ? csc_cb(r, g, b) : csc_cr(r, g, b);
}

// 24 bytes (8 RGB pixels) from p into six words, little-endian byte order.
template <int ALIGN>
__device__ __forceinline__ void load24(const uint8_t* p, uint32_t (&w)[6]) {
    if (ALIGN == 8) {
        const uint2* q = reinterpret_cast<const uint2*>(p);
        uint2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
        w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; w[5] = c.y;
    } else if (ALIGN == 4) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
        for (int j = 0; j < 6; ++j) w[j] = __ldg(q + j);
    } else {
#pragma unroll
        for (int j = 0; j < 6; ++j)
            w[j] = (uint32_t)__ldg(p + 4 * j) | ((uint32_t)__ldg(p + 4 * j + 1) << 8) |
                   ((uint32_t)__ldg(p + 4 * j + 2) << 16) | ((uint32_t)__ldg(p + 4 * j + 3) << 24);
    }
}

// getValueCategory / valueToBitString (utils.cpp:623-653): category = bit length of |v| (0 for v == 0); the value
// bits are v for v > 0 and v + 2^cat - 1 for v < 0, in `cat` bits.
__device__ __forceinline__ void cat_bits(int v, int& cat, uint32_t& vb) {
    cat = 32 - __clz(abs(v));
    vb = (uint32_t)(v + (v >> 31)) & ((1u << cat) - 1u);
}

// byte k (0..23, compile time) of the six words
__device__ __forceinline__ uint32_t byte24(const uint32_t (&w)[6], int k) { return (w[k >> 2] >> ((k & 3) * 8)) & 0xFFu; }

}  // namespace jb
