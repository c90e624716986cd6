// TEST INFRASTRUCTURE ONLY (oracle/).  Never linked into the product library.
//
// C-ABI harness around the reference's own CPU implementation.  It is compiled
// together with /root/reference/src/utils.cpp *unmodified* (see oracle/Makefile)
// into oracle/_ref/libjpegref.so.  Nothing from the reference is copied here:
// this file only calls the reference's functions (declared in its utils.hpp)
// in the order its driver does (src/OpenCLProject_JpegEncoder.cpp:59-225), with
// heap arrays instead of the driver's stack VLAs (cpp:190-191) and
// without the debug PPM dumps (cpp:32-123, untimed in the reference too).
//
// One extra transform is provided that is NOT in the reference:
// ref_dct_from_copy() evaluates the reference's DCT formula
// (src/utils.cpp:314-347, same summation order, same std::cos arguments, same
// final scale) but reads the 8x8 block from a private copy, i.e. without the
// in-place overwrite defect.  It is the "conformant" pin for the
// transform stage; everything else is the reference verbatim.
#include <OpenCL/cl-patched.hpp>

#include <chrono>
#include <cmath>
#include <cstring>
#include <pthread.h>
#include <string>
#include <vector>

#include "utils.hpp"    // the reference's header (-I/root/reference/src)
#include "huffman.hpp"  // the reference's code tables (for table export only)

namespace {

double now_us() {
    using namespace std::chrono;
    return duration<double, std::micro>(steady_clock::now().time_since_epoch()).count();
}

typedef const unsigned int (*qtab_t)[8];

// Out-of-place evaluation of the formula at src/utils.cpp:314-347.
void dct_block_from_copy(ppm_d_t *img, size_t x0, size_t y0) {
    rgb_pixel_d_t src[8][8];
    for (size_t j = 0; j < 8; ++j)
        for (size_t i = 0; i < 8; ++i) src[j][i] = img->data[(y0 + j) * img->width + (x0 + i)];
    for (size_t u = 0; u < 8; ++u) {
        for (size_t v = 0; v < 8; ++v) {
            double au = (u == 0) ? 1.0 / std::sqrt(2) : 1.0;
            double av = (v == 0) ? 1.0 / std::sqrt(2) : 1.0;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            for (size_t j = 0; j < 8; ++j) {
                for (size_t i = 0; i < 8; ++i) {
                    double cx = std::cos((2 * i + 1) * u * M_PI / 16.0);
                    double cy = std::cos((2 * j + 1) * v * M_PI / 16.0);
                    s0 += src[j][i].r * cx * cy;
                    s1 += src[j][i].g * cx * cy;
                    s2 += src[j][i].b * cx * cy;
                }
            }
            s0 *= (au * av / 4.0);
            s1 *= (au * av / 4.0);
            s2 *= (au * av / 4.0);
            rgb_pixel_d_t *dst = &img->data[(y0 + v) * img->width + (x0 + u)];
            dst->r = s0;
            dst->g = s1;
            dst->b = s2;
        }
    }
}

struct HuffArgs {
    int (*zz)[64];
    std::vector<std::vector<int>> *rle;
    int rpc;
    std::string out;
};

void *huff_thread(void *p) {
    HuffArgs *a = static_cast<HuffArgs *>(p);
    a->out = HuffmanEncoder(a->zz, *a->rle, a->rpc);
    return nullptr;
}

// HuffmanEncoder keeps `int dc_components[rpc][3]` on the stack (utils.cpp:661);
// run it on a thread whose stack is large enough for any rpc.
std::string huffman_big_stack(int (*zz)[64], std::vector<std::vector<int>> &rle, int rpc) {
    HuffArgs a{zz, &rle, rpc, std::string()};
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, (size_t)rpc * 12 + (64u << 20));
    pthread_t t;
    pthread_create(&t, &attr, huff_thread, &a);
    pthread_join(t, nullptr);
    pthread_attr_destroy(&attr);
    return a.out;
}

}  // namespace

extern "C" {

// Padded size exactly as the driver decides it (cpp:93-98).
void ref_padded_size(size_t W, size_t H, size_t *nW, size_t *nH) {
    if (W % 8 == 0 && H % 8 == 0) {
        *nW = W;
        *nH = H;
    } else {
        getNearest8x8ImageSize(W, H, nW, nH);
    }
}

void ref_performCSC(uint8_t *px, size_t W, size_t H) {
    ppm_t img{W, H, reinterpret_cast<rgb_pixel_t *>(px)};
    performCSC(&img);
}

void ref_performCDS(uint8_t *px, size_t W, size_t H) {
    ppm_t img{W, H, reinterpret_cast<rgb_pixel_t *>(px)};
    performCDS(&img);
}

void ref_pad(uint8_t *src, size_t W, size_t H, uint8_t *dst, size_t nW, size_t nH) {
    ppm_t a{W, H, reinterpret_cast<rgb_pixel_t *>(src)};
    ppm_t b{nW, nH, reinterpret_cast<rgb_pixel_t *>(dst)};
    copyToLargerImage(&a, &b);
    addReversedPadding(&b, W, H);
}

void ref_u8_to_double(uint8_t *src, double *dst, size_t W, size_t H) {
    ppm_t a{W, H, reinterpret_cast<rgb_pixel_t *>(src)};
    ppm_d_t b{W, H, reinterpret_cast<rgb_pixel_d_t *>(dst)};
    copyUIntToDoubleImage(&a, &b);
}

void ref_substractfromAll(double *img, size_t W, size_t H, double val) {
    ppm_d_t d{W, H, reinterpret_cast<rgb_pixel_d_t *>(img)};
    substractfromAll(&d, val);
}

// As written in the reference: in place.
void ref_performDCT(double *img, size_t W, size_t H) {
    ppm_d_t d{W, H, reinterpret_cast<rgb_pixel_d_t *>(img)};
    performDCT(&d);
}

// Same formula, block read from a copy (not in the reference).
void ref_dct_from_copy(double *img, size_t W, size_t H) {
    ppm_d_t d{W, H, reinterpret_cast<rgb_pixel_d_t *>(img)};
    for (size_t y = 0; y < H; y += 8)
        for (size_t x = 0; x < W; x += 8) dct_block_from_copy(&d, x, y);
}

void ref_performQuantization(double *img, size_t W, size_t H, const unsigned *ql, const unsigned *qc) {
    ppm_d_t d{W, H, reinterpret_cast<rgb_pixel_d_t *>(img)};
    performQuantization(&d, ql ? (qtab_t)ql : quant_mat_lum, qc ? (qtab_t)qc : quant_mat_chrom);
}

void ref_everyMCUisnow2DArray(double *img, size_t W, size_t H, int32_t *linear) {
    ppm_d_t d{W, H, reinterpret_cast<rgb_pixel_d_t *>(img)};
    everyMCUisnow2DArray(&d, reinterpret_cast<int(*)[64]>(linear));
}

void ref_performZigZag(int32_t *linear, int32_t *zz, int rows) {
    performZigZag(reinterpret_cast<int(*)[64]>(linear), reinterpret_cast<int(*)[64]>(zz), rows);
}

// performRLE over `rows` blocks; result flattened: counts[i] ints for block i.
uint64_t ref_performRLE(int32_t *zz, int rows, int32_t *flat, uint64_t cap, uint32_t *counts) {
    std::vector<std::vector<int>> rle;
    performRLE(reinterpret_cast<int(*)[64]>(zz), rle, rows);
    uint64_t n = 0;
    for (int i = 0; i < rows; ++i) {
        if (counts) counts[i] = (uint32_t)rle[i].size();
        for (int v : rle[i]) {
            if (flat && n < cap) flat[n] = v;
            ++n;
        }
    }
    return n;
}

// performRLE on all 3*rpc rows (cpp:213) then HuffmanEncoder (cpp:225).
// Returns the number of bits; writes up to cap '0'/'1' chars.
uint64_t ref_HuffmanEncoder(int32_t *zz, int rpc, char *bits, uint64_t cap) {
    std::vector<std::vector<int>> rle;
    performRLE(reinterpret_cast<int(*)[64]>(zz), rle, rpc * 3);
    std::string s = huffman_big_stack(reinterpret_cast<int(*)[64]>(zz), rle, rpc);
    if (bits) memcpy(bits, s.data(), s.size() < cap ? s.size() : cap);
    return s.size();
}

int ref_getValueCategory(int v) { return getValueCategory((int16_t)v); }

int ref_valueToBitString(int v, char *out) {
    std::string s = valueToBitString((int16_t)v);
    memcpy(out, s.data(), s.size());
    return (int)s.size();
}

// table: 0 DC luma, 1 DC chroma, 2 AC luma, 3 AC chroma.  Returns the string
// exactly as huffman.hpp holds it ("NULL" placeholders included).
const char *ref_table_code(int table, int run, int cat) {
    switch (table) {
        case 0: return DC_LUMA_HUFF_CODES[cat].c_str();
        case 1: return DC_CHROMA_HUFF_CODES[cat].c_str();
        case 2: return AC_LUMA_HUFF_CODES[run][cat].c_str();
        default: return AC_CHROMA_HUFF_CODES[run][cat].c_str();
    }
}

void ref_quant_tables(unsigned *ql, unsigned *qc) {
    memcpy(ql, quant_mat_lum, 64 * sizeof(unsigned));
    memcpy(qc, quant_mat_chrom, 64 * sizeof(unsigned));
}

// The whole CPU path in the driver's order with the driver's timer placement
// (cpp:58-247).  stage_us[9] mirrors CPUTelemetry (utils.hpp:65-75):
// CSC, CDS, levelShift, DCT, Quant, TotalCopy, zigZag, RLE, Huffman.
// dct_mode 0 = as written (in place), 1 = block read from a copy.
// rgb is not modified (the driver mutates its input; we work on a copy).
int ref_run_pipeline(const uint8_t *rgb, size_t W, size_t H, int dct_mode, const unsigned *ql,
                     const unsigned *qc, uint8_t *ycc_padded, int32_t *zigzag_out, char *bits,
                     uint64_t bits_cap, uint64_t *nbits, double *stage_us) {
    double t[9] = {0};
    std::vector<uint8_t> work(rgb, rgb + W * H * 3);
    ppm_t img{W, H, reinterpret_cast<rgb_pixel_t *>(work.data())};

    double t0 = now_us();
    performCSC(&img);
    t[0] = now_us() - t0;

    t0 = now_us();
    performCDS(&img);
    t[1] = now_us() - t0;

    size_t nW, nH;
    ref_padded_size(W, H, &nW, &nH);
    std::vector<uint8_t> padded(nW * nH * 3);
    ppm_t img3{nW, nH, reinterpret_cast<rgb_pixel_t *>(padded.data())};
    t0 = now_us();
    copyToLargerImage(&img, &img3);
    t[5] = now_us() - t0;
    addReversedPadding(&img3, W, H);  // untimed in the reference (cpp:120)
    if (ycc_padded) memcpy(ycc_padded, padded.data(), padded.size());

    std::vector<double> dbl(nW * nH * 3);
    ppm_d_t imgd{nW, nH, reinterpret_cast<rgb_pixel_d_t *>(dbl.data())};
    t0 = now_us();
    copyUIntToDoubleImage(&img3, &imgd);
    t[5] += now_us() - t0;

    t0 = now_us();
    substractfromAll(&imgd, 128.0);
    t[2] = now_us() - t0;

    t0 = now_us();
    if (dct_mode == 0)
        performDCT(&imgd);
    else
        ref_dct_from_copy(dbl.data(), nW, nH);
    t[3] = now_us() - t0;

    t0 = now_us();
    performQuantization(&imgd, ql ? (qtab_t)ql : quant_mat_lum, qc ? (qtab_t)qc : quant_mat_chrom);
    t[4] = now_us() - t0;

    unsigned rpc = (unsigned)(nW * nH / 64);
    unsigned rows = rpc * 3;
    std::vector<int32_t> lin((size_t)rows * 64), zz((size_t)rows * 64);
    t0 = now_us();
    everyMCUisnow2DArray(&imgd, reinterpret_cast<int(*)[64]>(lin.data()));
    performZigZag(reinterpret_cast<int(*)[64]>(lin.data()), reinterpret_cast<int(*)[64]>(zz.data()), rows);
    t[6] = now_us() - t0;
    if (zigzag_out) memcpy(zigzag_out, zz.data(), zz.size() * sizeof(int32_t));

    std::vector<std::vector<int>> rle;
    t0 = now_us();
    performRLE(reinterpret_cast<int(*)[64]>(zz.data()), rle, rows);
    t[7] = now_us() - t0;

    t0 = now_us();
    std::string s = huffman_big_stack(reinterpret_cast<int(*)[64]>(zz.data()), rle, (int)rpc);
    t[8] = now_us() - t0;

    if (nbits) *nbits = s.size();
    if (bits) memcpy(bits, s.data(), s.size() < bits_cap ? s.size() : bits_cap);
    if (stage_us) memcpy(stage_us, t, sizeof(t));
    return 0;
}

// The OpenCL half's host helpers (utils.hpp:116,119), called as main() calls them (cpp:325).
void ref_copyImageToVector(uint8_t *px, size_t W, size_t H, uint32_t *v) {
    ppm_t img;
    img.width = W;
    img.height = H;
    img.data = reinterpret_cast<rgb_pixel_t *>(px);
    std::vector<cl_uint> vec(W * H * 3);
    copyImageToVector(&img, vec);
    memcpy(v, vec.data(), vec.size() * sizeof(cl_uint));
}
void ref_switchVectorChannelOrdering(const uint32_t *in, size_t W, size_t H, uint32_t *out) {
    std::vector<cl_uint> a(in, in + W * H * 3), b(W * H * 3);
    switchVectorChannelOrdering(a, b, (unsigned)W, (unsigned)H);
    memcpy(out, b.data(), b.size() * sizeof(cl_uint));
}

// The remaining helpers of utils.hpp that the product mirrors (utils.hpp:79, 95, 118, 123), called verbatim.
void ref_removeRedChannel(uint8_t *px, size_t W, size_t H) {
    ppm_t img{W, H, reinterpret_cast<rgb_pixel_t *>(px)};
    removeRedChannel(&img);
}
void ref_copyDoubleToUIntImage(double *src, size_t W, size_t H, uint8_t *dst) {
    ppm_d_t a{W, H, reinterpret_cast<rgb_pixel_d_t *>(src)};
    ppm_t b{W, H, reinterpret_cast<rgb_pixel_t *>(dst)};
    copyDoubleToUIntImage(&a, &b);
}
// The reference's bottom loop reads vInput[(oldH - diff) * oldW + x] with x up to newW (utils.cpp:733-740): past
// the row, and for the last plane past the vector.  The input is therefore handed over with newW spare words at
// its end so that the call is defined; the corner region it fills is not compared by the tests.
void ref_copyOntoLargerVectorWithPadding(const uint32_t *in, size_t W, size_t H, uint32_t *out, size_t nW, size_t nH) {
    std::vector<cl_uint> a(in, in + W * H * 3), b(nW * nH * 3);
    a.resize(W * H * 3 + nW + 8, 0);
    copyOntoLargerVectorWithPadding(a, b, (unsigned)W, (unsigned)H, (unsigned)nW, (unsigned)nH);
    memcpy(out, b.data(), b.size() * sizeof(cl_uint));
}
void ref_everyMCUisnow1DArray(const int32_t *in, size_t W, size_t H, int32_t *out) {
    std::vector<int> a(in, in + W * H * 3);
    everyMCUisnow1DArray(a, out, (unsigned)W, (unsigned)H);
}

}  // extern "C"
