// utils_compat.hpp -- the reference's per-stage function surface (src/utils.hpp:7-39,
// 77-137 of rusty-electron/jpeg-encoder-opencl) re-declared on top of the C ABI of
// libjpegb200.so, so that a driver written against the reference (JpegEncoderHost,
// src/OpenCLProject_JpegEncoder.cpp:28-250) can be re-linked against the B200 path
// unchanged: same names, same argument types, same in-place/out-parameter behaviour.
// std::vector / std::string never cross the C ABI; this shim converts.
//
// Differences a caller can observe (all documented in DESIGN.md):
//  * performDCT computes the true DCT.  Set jb_compat::flags() |= JB_FLAG_REF_INPLACE_DCT
//    to reproduce the reference's in-place overwrite bit for bit.
//  * HuffmanEncoder emits Annex-K codes and suppresses EOB after a full block unless
//    JB_FLAG_REF_TYPO_TABLES / JB_FLAG_REF_ALWAYS_EOB are set.
//  * errors throw std::runtime_error (the reference's stage functions have no error path).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "jpegb200.h"

struct rgb_pixel { uint8_t r, g, b; };                 // utils.hpp:7-11 (after CSC: Y, Cb, Cr)
typedef struct rgb_pixel rgb_pixel_t;
struct PPMimage { size_t width, height; rgb_pixel_t* data; };   // utils.hpp:15-19
typedef struct PPMimage ppm_t;
struct rgb_pixel_d { double r, g, b; };                // utils.hpp:25-29
typedef struct rgb_pixel_d rgb_pixel_d_t;
struct PPMimage_d { size_t width, height; rgb_pixel_d_t* data; };  // utils.hpp:33-37
typedef struct PPMimage_d ppm_d_t;

// utils.hpp:41-62 (T.81 Annex K.1 / K.2)
static const unsigned int quant_mat_lum[8][8] = {
    {16, 11, 10, 16, 24, 40, 51, 61},     {12, 12, 14, 19, 26, 58, 60, 55},
    {14, 13, 16, 24, 40, 57, 69, 56},     {14, 17, 22, 29, 51, 87, 80, 62},
    {18, 22, 37, 56, 68, 109, 103, 77},   {24, 35, 55, 64, 81, 104, 113, 92},
    {49, 64, 78, 87, 103, 121, 120, 101}, {72, 92, 95, 98, 112, 100, 103, 99}};
static const unsigned int quant_mat_chrom[8][8] = {
    {17, 18, 24, 47, 99, 99, 99, 99}, {18, 21, 26, 66, 99, 99, 99, 99}, {24, 26, 56, 99, 99, 99, 99, 99},
    {47, 66, 99, 99, 99, 99, 99, 99}, {99, 99, 99, 99, 99, 99, 99, 99}, {99, 99, 99, 99, 99, 99, 99, 99},
    {99, 99, 99, 99, 99, 99, 99, 99}, {99, 99, 99, 99, 99, 99, 99, 99}};

namespace jb_compat {
inline jb_ctx*& ctx_slot() { static jb_ctx* c = nullptr; return c; }
inline uint32_t& flags() { static uint32_t f = 0; return f; }
inline jb_ctx* ctx() {
    jb_ctx*& c = ctx_slot();
    if (!c && jb_create(0, &c) != JB_OK) throw std::runtime_error("jb_create failed: no CUDA device (no CPU fallback)");
    return c;
}
inline void ck(int rc) { if (rc != JB_OK) throw std::runtime_error(jb_last_error(ctx())); }
inline void shutdown() { if (ctx_slot()) { jb_destroy(ctx_slot()); ctx_slot() = nullptr; } }
}  // namespace jb_compat

// ---- I/O (utils.cpp:11-82): binary P6, '#' comments, maxval 255; 0 / -1 ---------------
inline int readPPMImage(const char* path, size_t* width, size_t* height, rgb_pixel_t** imgptr) {
    FILE* fp = fopen(path, "rb");
    if (!fp) { printf("Error opening the file\n"); return -1; }
    char line[128];
    if (!fgets(line, sizeof line, fp) || strcmp(line, "P6\n")) { printf("Invalid file format\n"); fclose(fp); return -1; }
    do { if (!fgets(line, sizeof line, fp)) { fclose(fp); return -1; } } while (line[0] == '#');
    unsigned long w = 0, h = 0;
    if (sscanf(line, "%lu %lu", &w, &h) != 2) { fclose(fp); return -1; }
    if (!fgets(line, sizeof line, fp) || atoi(line) != 255) { printf("Invalid maximum value\n"); fclose(fp); return -1; }
    *width = w; *height = h;
    *imgptr = (rgb_pixel_t*)malloc(w * h * sizeof(rgb_pixel_t));
    if (!*imgptr) { printf("Error allocating memory\n"); fclose(fp); return -1; }
    size_t got = fread(*imgptr, sizeof(rgb_pixel_t), w * h, fp);
    fclose(fp);
    return got == w * h ? 0 : -1;
}
inline int writePPMImage(const char* path, size_t width, size_t height, rgb_pixel_t* img) {
    FILE* fp = fopen(path, "wb");
    if (!fp) { printf("Error opening the file\n"); return -1; }
    fprintf(fp, "P6\n%zu %zu\n255\n", width, height);
    fwrite(img, sizeof(rgb_pixel_t), width * height, fp);
    fclose(fp);
    return 0;
}

// ---- stages --------------------------------------------------------------------------
inline void performCSC(ppm_t* img) {                                           // utils.hpp:81
    jb_compat::ck(jb_csc_rgb8_aos(jb_compat::ctx(), (uint8_t*)img->data, img->width, img->height));
}
inline void performCDS(ppm_t* img) {                                           // utils.hpp:82
    jb_compat::ck(jb_cds_aos(jb_compat::ctx(), (uint8_t*)img->data, img->width, img->height));
}
inline void getNearest8x8ImageSize(size_t w, size_t h, size_t* nw, size_t* nh) {  // utils.hpp:98
    jb_padded_size(w, h, 8, nw, nh);
}
// utils.hpp:97 + 99.  The reference first copies, then mirrors in place; here both calls
// produce the complete padded image (copying is a subset of it), so either order works.
inline void copyToLargerImage(ppm_t* img, ppm_t* newImg) {
    jb_compat::ck(jb_pad_mirror_aos(jb_compat::ctx(), (const uint8_t*)img->data, img->width, img->height,
                                    (uint8_t*)newImg->data, newImg->width, newImg->height));
}
inline void addReversedPadding(ppm_t* img, size_t oldWidth, size_t oldHeight) {
    std::vector<rgb_pixel_t> tight(oldWidth * oldHeight);  // marshal the top-left region to a tight buffer
    for (size_t y = 0; y < oldHeight; ++y) memcpy(&tight[y * oldWidth], &img->data[y * img->width], oldWidth * 3);
    jb_compat::ck(jb_pad_mirror_aos(jb_compat::ctx(), (const uint8_t*)tight.data(), oldWidth, oldHeight,
                                    (uint8_t*)img->data, img->width, img->height));
}
inline void copyUIntToDoubleImage(ppm_t* img, ppm_d_t* newImg) {               // utils.hpp:94
    jb_compat::ck(jb_u8_to_f64(jb_compat::ctx(), (const uint8_t*)img->data, (double*)newImg->data,
                               img->width * img->height * 3));
}
inline void substractfromAll(ppm_d_t* img, double val) {                       // utils.hpp:100
    jb_compat::ck(jb_levelshift_f64(jb_compat::ctx(), (double*)img->data, img->width * img->height * 3, val));
}
inline void performDCT(ppm_d_t* img) {                                         // utils.hpp:102
    jb_compat::ck(jb_dct_f64(jb_compat::ctx(), (double*)img->data, img->width, img->height, jb_compat::flags()));
}
inline void performQuantization(ppm_d_t* img, const unsigned int ql[][8], const unsigned int qc[][8]) {  // :106
    jb_compat::ck(jb_quantize_f64(jb_compat::ctx(), (double*)img->data, img->width, img->height,
                                  (const uint32_t*)ql, (const uint32_t*)qc));
}
inline void everyMCUisnow2DArray(ppm_d_t* img, int linear_arr[][64]) {         // utils.hpp:122
    jb_compat::ck(jb_blockify(jb_compat::ctx(), (const double*)img->data, img->width, img->height,
                              (int32_t*)linear_arr));
}
inline void performZigZag(int linear_arr[][64], int zigzag_arr[][64], int numRows) {  // utils.hpp:127
    jb_compat::ck(jb_zigzag(jb_compat::ctx(), (const int32_t*)linear_arr, (int32_t*)zigzag_arr, (size_t)numRows));
}
// ---- the planar uint32 image of the reference's OpenCL half (cl_uint = uint32_t) -------------------
typedef uint32_t cl_uint;
inline void copyImageToVector(ppm_t* img, std::vector<cl_uint>& v) {           // utils.hpp:116
    jb_compat::ck(jb_planar_u32_from_aos(jb_compat::ctx(), (const uint8_t*)img->data, img->width, img->height, v.data()));
}
inline void switchVectorChannelOrdering(std::vector<cl_uint>& vInput, std::vector<cl_uint>& vOutput,
                                        const unsigned int width, const unsigned int height) {  // utils.hpp:119
    jb_compat::ck(jb_planar_u32_interleave(jb_compat::ctx(), vInput.data(), width, height, vOutput.data()));
}
inline void performRLE(int zigzag_array[][64], std::vector<std::vector<int>>& rle, int rows) {  // utils.hpp:132
    std::vector<int32_t> pairs((size_t)rows * 128);
    std::vector<uint32_t> counts((size_t)rows);
    jb_compat::ck(jb_rle(jb_compat::ctx(), (const int32_t*)zigzag_array, (size_t)rows, jb_compat::flags(),
                         pairs.data(), counts.data()));
    for (int i = 0; i < rows; ++i)
        rle.emplace_back(pairs.begin() + (size_t)i * 128, pairs.begin() + (size_t)i * 128 + counts[i]);
}
// utils.hpp:137.  The rle argument is accepted for signature compatibility; the run-lengths are
// recomputed on the GPU from zigzag_array (they are a pure function of it, utils.cpp:612-620).
inline std::string HuffmanEncoder(int zigzag_array[][64], std::vector<std::vector<int>>& /*rle*/, int rowsPerChannel) {
    size_t cap = (size_t)rowsPerChannel * 3 * 64 * 4 + 64;
    std::vector<uint8_t> packed(cap);
    uint64_t nbits = 0;
    jb_compat::ck(jb_huffman(jb_compat::ctx(), (const int32_t*)zigzag_array, (size_t)rowsPerChannel,
                             jb_compat::flags(), packed.data(), cap, &nbits));
    std::string s((size_t)nbits, '0');
    for (uint64_t i = 0; i < nbits; ++i)
        if (packed[i >> 3] & (0x80u >> (i & 7))) s[(size_t)i] = '1';
    return s;
}
