// Per-stage kernels behind the staged C-ABI entry points: one launch per
// reference function (src/utils.hpp:81-137) on the reference's own layouts, so
// that each stage of JpegEncoderHost (src/OpenCLProject_JpegEncoder.cpp:59-225)
// has a drop-in.  They share the arithmetic of the fused kernel (jb_math.h,
// jb_pixels.cuh) and, for the binary64 stages, repeat the reference's operation
// order so that results are bit-identical.  Also: the synthetic image generator.
#include "jb_pixels.cuh"

namespace jb {

static inline int grid_for(size_t n, int block) {
    size_t g = (n + block - 1) / block;
    return (int)(g < 1 ? 1 : g > 148 * 64 ? 148 * 64 : g);
}

// performCSC, utils.cpp:92-110 (in place on AoS bytes)
__global__ void k_csc(uint8_t* px, size_t n, const uint32_t* ydown) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t r = px[3 * i], g = px[3 * i + 1], b = px[3 * i + 2];
        px[3 * i] = (uint8_t)csc_y(r, g, b, ydown);
        px[3 * i + 1] = (uint8_t)csc_cb(r, g, b);
        px[3 * i + 2] = (uint8_t)csc_cr(r, g, b);
    }
}

// performCDS, utils.cpp:113-141: one thread per complete 2x2 cell
__global__ void k_cds(uint8_t* px, size_t W, size_t H) {
    size_t cw = W / 2, ch = H / 2, n = cw * ch;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t x = (i % cw) * 2, y = (i / cw) * 2;
        uint8_t* p[4] = {px + 3 * (y * W + x), px + 3 * (y * W + x + 1), px + 3 * ((y + 1) * W + x),
                         px + 3 * ((y + 1) * W + x + 1)};
#pragma unroll
        for (int c = 1; c <= 2; ++c) {
            uint32_t m = ((uint32_t)p[0][c] + p[1][c] + p[2][c] + p[3][c]) >> 2;
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k][c] = (uint8_t)m;
        }
    }
}

// copyToLargerImage + addReversedPadding, utils.cpp:199-233
__global__ void k_pad(const uint8_t* src, size_t W, size_t H, uint8_t* dst, size_t nW, size_t nH) {
    size_t n = nW * nH;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int x = (int)(i % nW), y = (int)(i / nW);
        const uint8_t* s = src + 3 * ((size_t)mirror(y, (int)H) * W + (size_t)mirror(x, (int)W));
        dst[3 * i] = s[0];
        dst[3 * i + 1] = s[1];
        dst[3 * i + 2] = s[2];
    }
}

__global__ void k_u8_to_f64(const uint8_t* src, double* dst, size_t n) {  // utils.cpp:236-246
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (double)src[i];
}

__global__ void k_sub_f64(double* img, size_t n, double val) {  // utils.cpp:190-196
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        img[i] = __dsub_rn(img[i], val);
}

// performDCT / performDCTBlock, utils.cpp:262-270, 314-347.  One thread per
// (block, channel); operation order of the reference: outputs u outer / v inner,
// sum y outer / x inner, term (sample * cos_x) * cos_y, scale last.  inplace
// keeps the reference's overwrite-while-reading behaviour.
__global__ void k_dct_f64(double* img, size_t W, size_t H, int inplace, const double* __restrict__ costab,
                          const double* __restrict__ scale) {
    size_t bx = W / 8, nb = bx * (H / 8) * 3;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nb; t += (size_t)gridDim.x * blockDim.x) {
        size_t blk = t / 3, c = t % 3;
        size_t x0 = (blk % bx) * 8, y0 = (blk / bx) * 8;
        double in[64], out[64];
        for (int j = 0; j < 8; ++j)
            for (int i = 0; i < 8; ++i) in[j * 8 + i] = img[3 * ((y0 + j) * W + x0 + i) + c];
        double* dstv = inplace ? in : out;
        for (int u = 0; u < 8; ++u)
            for (int v = 0; v < 8; ++v) {
                double s = 0.0;
                for (int y = 0; y < 8; ++y)
                    for (int x = 0; x < 8; ++x)
                        s = __dadd_rn(s, __dmul_rn(__dmul_rn(in[y * 8 + x], costab[u * 8 + x]), costab[v * 8 + y]));
                dstv[v * 8 + u] = __dmul_rn(s, scale[u * 8 + v]);
            }
        for (int j = 0; j < 8; ++j)
            for (int i = 0; i < 8; ++i) img[3 * ((y0 + j) * W + x0 + i) + c] = dstv[j * 8 + i];
    }
}

// performQuantization, utils.cpp:454-467
__global__ void k_quant_f64(double* img, size_t W, size_t H, const __grid_constant__ QuantTables qt) {
    size_t n = W * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t x = i % W, y = i / W;
        int k = (int)((y & 7) * 8 + (x & 7));
        img[3 * i] = round(__ddiv_rn(img[3 * i], (double)qt.q[0][k]));
        img[3 * i + 1] = round(__ddiv_rn(img[3 * i + 1], (double)qt.q[1][k]));
        img[3 * i + 2] = round(__ddiv_rn(img[3 * i + 2], (double)qt.q[1][k]));
    }
}

// everyMCUisnow2DArray, utils.cpp:482-498
__global__ void k_blockify(const double* img, size_t W, size_t H, int32_t* linear) {
    size_t n = W * H, rpc = n / 64, bx = W / 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t x = i % W, y = i / W;
        size_t blk = (y / 8) * bx + x / 8, k = (y & 7) * 8 + (x & 7);
#pragma unroll
        for (int c = 0; c < 3; ++c) linear[(blk + rpc * c) * 64 + k] = (int)img[3 * i + c];
    }
}

// performZigZag, utils.cpp:539-558
__global__ void k_zigzag(const int32_t* linear, int32_t* zz, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        zz[i] = linear[(i & ~(size_t)63) + c_zz[i & 63]];
}

// performRLE / RLEBlockAC, utils.cpp:572-620: one thread per block
__global__ void k_rle(const int32_t* zz, size_t rows, int always_eob, int32_t* pairs, uint32_t* counts) {
    for (size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (size_t)gridDim.x * blockDim.x) {
        const int32_t* z = zz + r * 64;
        int32_t* o = pairs + r * 128;
        int last = 0;
        for (int i = 63; i >= 0; --i)
            if (z[i] != 0) {
                last = i;
                break;
            }
        uint32_t n = 0;
        int run = 0;
        for (int i = 1; i <= last; ++i) {
            if (z[i] == 0) {
                if (run == 15) {
                    o[n++] = 15;
                    o[n++] = 0;
                    run = 0;
                } else {
                    ++run;
                }
            } else {
                o[n++] = run;
                o[n++] = z[i];
                run = 0;
            }
        }
        if (always_eob || last != 63) {
            o[n++] = 0;
            o[n++] = 0;
        }
        counts[r] = n;
    }
}

// reference layout int32[3*rpc][64] (planar by channel) -> scan order int16[rpc][3][64]
__global__ void k_planar_to_scan(const int32_t* zz, size_t rpc, int16_t* coef) {
    size_t n = rpc * 3 * 64;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t k = i & 63, b = i >> 6, m = b / 3, c = b % 3;
        coef[i] = (int16_t)zz[(m + rpc * c) * 64 + k];
    }
}

// Synthetic image of SURVEY.md section 8d (integer only, so that the CPU checker can regenerate it bit for bit)
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void k_synth(uint64_t seed, size_t W, size_t y0, size_t rows, size_t pitch, uint8_t* out) {
    size_t n = W * rows;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        size_t x = i % W, y = y0 + i / W;
        uint8_t* o = out + (i / W) * pitch + x * 3;
#pragma unroll
        for (unsigned c = 0; c < 3; ++c) {
            const unsigned P = c == 0 ? 97u : c == 1 ? 61u : 41u, A = c == 2 ? 2u : 1u, B = c == 1 ? 2u : 1u;
            unsigned ph = (unsigned)((x * A + y * B) % P);
            unsigned v = ph <= P / 2 ? ph : P - ph;
            int base = 32 + (int)(v * 192 / (P / 2));
            uint64_t h = splitmix64(seed ^ ((((uint64_t)y << 32) | (uint64_t)x) * 3 + c));
            int val = base + (int)((h >> 56) % 9) - 4;
            o[c] = (uint8_t)(val < 0 ? 0 : val > 255 ? 255 : val);
        }
    }
}

// ---- the planar uint32 image layout of the reference's OpenCL half ----------------------------
// copyImageToVector, utils.cpp:700-707: AoS bytes -> R plane, G plane, B plane of n = W*H words each
__global__ void k_aos_to_planar_u32(const uint8_t* px, size_t n, uint32_t* out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        out[i] = px[3 * i];
        out[i + n] = px[3 * i + 1];
        out[i + 2 * n] = px[3 * i + 2];
    }
}
// switchVectorChannelOrdering, utils.cpp:745-754: planes -> interleaved words (RGBRGB...)
__global__ void k_planar_u32_interleave(const uint32_t* in, size_t n, uint32_t* out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        out[3 * i] = in[i];
        out[3 * i + 1] = in[i + n];
        out[3 * i + 2] = in[i + 2 * n];
    }
}
// planes (as the reference's kernels index them, .cl:17-19: d_input[c*W*H + y*W + x]) -> the pitched AoS RGB8
// frame the fused kernels read; the low byte of every word is the sample (the reference stores bytes in words)
__global__ void k_planar_u32_to_rgb8(const uint32_t* in, size_t W, size_t H, uint8_t* out, size_t pitch) {
    const size_t n = W * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t y = i / W, x = i - y * W;
        uint8_t* o = out + y * pitch + 3 * x;
        o[0] = (uint8_t)in[i];
        o[1] = (uint8_t)in[i + n];
        o[2] = (uint8_t)in[i + 2 * n];
    }
}

// copyOntoLargerVectorWithPadding, utils.cpp:710-741: the mirror padding of addReversedPadding on the planar
// uint32 image of the OpenCL half.  (The reference's bottom loop indexes the *unpadded* input with x up to the new
// width, i.e. it reads the next row -- and past the vector for the last plane -- in the bottom-right corner; here
// the corner mirrors in both directions like the CPU path, utils.cpp:223-232.)
__global__ void k_pad_planar_u32(const uint32_t* in, size_t W, size_t H, uint32_t* out, size_t nW, size_t nH) {
    const size_t n = nW * nH;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % nW), y = (int)(i / nW);
        const size_t s = (size_t)mirror(y, (int)H) * W + (size_t)mirror(x, (int)W);
#pragma unroll
        for (int c = 0; c < 3; ++c) out[i + c * n] = in[s + c * W * H];
    }
}

// everyMCUisnow1DArray, utils.cpp:501-515: planar int image (plane c at c*W*H) -> int[3*rpc][64], blocks in raster
// order, planar by channel -- the block array of everyMCUisnow2DArray built from the OpenCL half's layout
__global__ void k_blockify_planar_i32(const int32_t* in, size_t W, size_t H, int32_t* linear) {
    const size_t n = W * H, rpc = n / 64, bx = W / 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t x = i % W, y = i / W;
        const size_t blk = (y / 8) * bx + x / 8, k = (y & 7) * 8 + (x & 7);
#pragma unroll
        for (int c = 0; c < 3; ++c) linear[(blk + rpc * c) * 64 + k] = in[i + c * n];
    }
}

// copyDoubleToUIntImage, utils.cpp:249-259: (uint8_t) of a double (truncation; out-of-range values wrap like the
// reference's x86-64 build: cvttsd2si, low byte)
__global__ void k_f64_to_u8(const double* src, uint8_t* dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (uint8_t)(uint32_t)__double2int_rz(src[i]);
}

// getValueCategory / valueToBitString, utils.cpp:623-653, for n values: the very cat_bits() the entropy coder uses
__global__ void k_value_categories(const int16_t* v, size_t n, uint8_t* cat, uint16_t* bits) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        int c;
        uint32_t vb;
        cat_bits((int)v[i], c, vb);
        cat[i] = (uint8_t)c;
        bits[i] = (uint16_t)vb;
    }
}

// removeRedChannel, utils.cpp:84-89 (the reference's "TEST FUNCTION"): r = 0 for every pixel
__global__ void k_remove_red(uint8_t* px, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) px[3 * i] = 0;
}

int launch_pad_planar_u32(const uint32_t* in, size_t W, size_t H, uint32_t* out, size_t nW, size_t nH, cudaStream_t s) {
    k_pad_planar_u32<<<grid_for(nW * nH, 256), 256, 0, s>>>(in, W, H, out, nW, nH);
    return 1;
}
int launch_blockify_planar_i32(const int32_t* in, size_t W, size_t H, int32_t* linear, cudaStream_t s) {
    k_blockify_planar_i32<<<grid_for(W * H, 256), 256, 0, s>>>(in, W, H, linear);
    return 1;
}
int launch_f64_to_u8(const double* src, uint8_t* dst, size_t n, cudaStream_t s) {
    k_f64_to_u8<<<grid_for(n, 256), 256, 0, s>>>(src, dst, n);
    return 1;
}
int launch_value_categories(const int16_t* v, size_t n, uint8_t* cat, uint16_t* bits, cudaStream_t s) {
    k_value_categories<<<grid_for(n, 256), 256, 0, s>>>(v, n, cat, bits);
    return 1;
}
int launch_remove_red(uint8_t* px, size_t n, cudaStream_t s) {
    k_remove_red<<<grid_for(n, 256), 256, 0, s>>>(px, n);
    return 1;
}

int launch_csc(uint8_t* px, size_t n, const uint32_t* ydown, cudaStream_t s) {
    k_csc<<<grid_for(n, 256), 256, 0, s>>>(px, n, ydown);
    return 1;
}
int launch_cds(uint8_t* px, size_t W, size_t H, cudaStream_t s) {
    k_cds<<<grid_for((W / 2) * (H / 2), 256), 256, 0, s>>>(px, W, H);
    return 1;
}
int launch_pad(const uint8_t* src, size_t W, size_t H, uint8_t* dst, size_t nW, size_t nH, cudaStream_t s) {
    k_pad<<<grid_for(nW * nH, 256), 256, 0, s>>>(src, W, H, dst, nW, nH);
    return 1;
}
int launch_u8_to_f64(const uint8_t* src, double* dst, size_t n, cudaStream_t s) {
    k_u8_to_f64<<<grid_for(n, 256), 256, 0, s>>>(src, dst, n);
    return 1;
}
int launch_sub_f64(double* img, size_t n, double val, cudaStream_t s) {
    k_sub_f64<<<grid_for(n, 256), 256, 0, s>>>(img, n, val);
    return 1;
}
int launch_dct_f64(double* img, size_t W, size_t H, int inplace, const double* costab, const double* scale,
                   cudaStream_t s) {
    k_dct_f64<<<grid_for((W / 8) * (H / 8) * 3, 64), 64, 0, s>>>(img, W, H, inplace, costab, scale);
    return 1;
}
int launch_quant_f64(double* img, size_t W, size_t H, const QuantTables& qt, cudaStream_t s) {
    k_quant_f64<<<grid_for(W * H, 256), 256, 0, s>>>(img, W, H, qt);
    return 1;
}
int launch_blockify(const double* img, size_t W, size_t H, int32_t* linear, cudaStream_t s) {
    k_blockify<<<grid_for(W * H, 256), 256, 0, s>>>(img, W, H, linear);
    return 1;
}
int launch_zigzag(const int32_t* linear, int32_t* zz, size_t rows, cudaStream_t s) {
    k_zigzag<<<grid_for(rows * 64, 256), 256, 0, s>>>(linear, zz, rows * 64);
    return 1;
}
int launch_rle(const int32_t* zz, size_t rows, int always_eob, int32_t* pairs, uint32_t* counts, cudaStream_t s) {
    k_rle<<<grid_for(rows, 128), 128, 0, s>>>(zz, rows, always_eob, pairs, counts);
    return 1;
}
int launch_planar_to_scan(const int32_t* zz, size_t rpc, int16_t* coef, cudaStream_t s) {
    k_planar_to_scan<<<grid_for(rpc * 192, 256), 256, 0, s>>>(zz, rpc, coef);
    return 1;
}
int launch_synth(uint64_t seed, size_t W, size_t y0, size_t rows, size_t pitch, uint8_t* d_out, cudaStream_t s) {
    k_synth<<<grid_for(W * rows, 256), 256, 0, s>>>(seed, W, y0, rows, pitch, d_out);
    return 1;
}

int launch_aos_to_planar_u32(const uint8_t* px, size_t n, uint32_t* out, cudaStream_t s) {
    k_aos_to_planar_u32<<<grid_for(n, 256), 256, 0, s>>>(px, n, out);
    return 1;
}
int launch_planar_u32_interleave(const uint32_t* in, size_t n, uint32_t* out, cudaStream_t s) {
    k_planar_u32_interleave<<<grid_for(n, 256), 256, 0, s>>>(in, n, out);
    return 1;
}
int launch_planar_u32_to_rgb8(const uint32_t* in, size_t W, size_t H, uint8_t* out, size_t pitch, cudaStream_t s) {
    k_planar_u32_to_rgb8<<<grid_for(W * H, 256), 256, 0, s>>>(in, W, H, out, pitch);
    return 1;
}

}  // namespace jb
