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
//
// Every name utils.hpp declares is here (tests/test_abi.py compiles the reference's own JpegEncoderHost,
// cpp:28-250, against this header).  Stage functions run on the GPU through the C ABI; the pixel accessors,
// the preview printers and the file writers are host-side conveniences of the reference's driver and stay
// host code here (they touch one pixel or print to stdout -- there is nothing to run on a device).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
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

// utils.hpp:65-75: per-stage times in microseconds -- the reference's only reporting surface.  jb_timings
// (jpegb200.h) starts with the same nine fields, so a jb_timings* can be read as a CPUTelemetry*.
struct CPUTelemetry {
    double CSCTime, CDSTime, levelShiftTime, DCTTime, QuantTime, TotalCopyTime, zigZagTime, RLETime, HuffmanTime;
};

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

// ---- pixel accessors (utils.hpp:84-92), host side ---------------------------------------------------------
inline rgb_pixel_t* getPixelPtr(ppm_t* img, size_t x, size_t y) { return img->data + y * img->width + x; }
inline rgb_pixel_t getPixel(ppm_t* img, size_t x, size_t y) { return *getPixelPtr(img, x, y); }
inline uint8_t getPixelR(ppm_t* img, size_t x, size_t y) { return getPixelPtr(img, x, y)->r; }
inline uint8_t getPixelG(ppm_t* img, size_t x, size_t y) { return getPixelPtr(img, x, y)->g; }
inline uint8_t getPixelB(ppm_t* img, size_t x, size_t y) { return getPixelPtr(img, x, y)->b; }
inline void setPixelR(ppm_t* img, size_t x, size_t y, uint8_t v) { getPixelPtr(img, x, y)->r = v; }
inline void setPixelG(ppm_t* img, size_t x, size_t y, uint8_t v) { getPixelPtr(img, x, y)->g = v; }
inline void setPixelB(ppm_t* img, size_t x, size_t y, uint8_t v) { getPixelPtr(img, x, y)->b = v; }

// ---- stages --------------------------------------------------------------------------
inline void removeRedChannel(ppm_t* img) {                                     // utils.hpp:79 ("TEST FUNCTION")
    jb_compat::ck(jb_remove_red_aos(jb_compat::ctx(), (uint8_t*)img->data, img->width, img->height));
}
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
inline void copyDoubleToUIntImage(ppm_d_t* img, ppm_t* newImg) {               // utils.hpp:95
    jb_compat::ck(jb_f64_to_u8(jb_compat::ctx(), (const double*)img->data, (uint8_t*)newImg->data,
                               img->width * img->height * 3));
}
inline void substractfromAll(ppm_d_t* img, double val) {                       // utils.hpp:100
    jb_compat::ck(jb_levelshift_f64(jb_compat::ctx(), (double*)img->data, img->width * img->height * 3, val));
}
inline void performDCT(ppm_d_t* img) {                                         // utils.hpp:102
    jb_compat::ck(jb_dct_f64(jb_compat::ctx(), (double*)img->data, img->width, img->height, jb_compat::flags()));
}
// utils.hpp:103-104.  performDCT2 is the reference's unused variant of the same transform (utils.cpp:273-311);
// performDCTBlock transforms one block in place (utils.cpp:314-347) -- here the block is marshalled to an 8x8
// image and goes through the same GPU stage.
inline void performDCT2(ppm_d_t* img) { performDCT(img); }
inline void performDCTBlock(ppm_d_t* img, size_t startX, size_t startY) {
    rgb_pixel_d_t blk[64];
    for (size_t y = 0; y < 8; ++y) memcpy(&blk[y * 8], &img->data[(startY + y) * img->width + startX], 8 * sizeof(rgb_pixel_d_t));
    ppm_d_t one{8, 8, blk};
    performDCT(&one);
    for (size_t y = 0; y < 8; ++y) memcpy(&img->data[(startY + y) * img->width + startX], &blk[y * 8], 8 * sizeof(rgb_pixel_d_t));
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
inline void copyOntoLargerVectorWithPadding(std::vector<cl_uint>& vInput, std::vector<cl_uint>& vOutput,
                                            const unsigned int oldWidth, const unsigned int oldHeight,
                                            const unsigned int newWidth, const unsigned int newHeight) {  // utils.hpp:118
    jb_compat::ck(jb_pad_mirror_planar_u32(jb_compat::ctx(), vInput.data(), oldWidth, oldHeight, vOutput.data(), newWidth,
                                           newHeight));
}
inline void writeVectorToFile(const char* path, const unsigned int width, const unsigned int height,
                              std::vector<cl_uint>& imgVector) {                 // utils.hpp:120: the words' low bytes as a P6
    FILE* fp = fopen(path, "wb");
    if (!fp) { printf("Error opening the file\n"); return; }
    fprintf(fp, "P6\n%u %u\n255\n", width, height);
    std::vector<uint8_t> bytes((size_t)3 * width * height);
    for (size_t i = 0; i < bytes.size(); ++i) bytes[i] = (uint8_t)imgVector[i];
    fwrite(bytes.data(), 1, bytes.size(), fp);
    fclose(fp);
}
inline void everyMCUisnow1DArray(std::vector<int>& input_arr, int output_arr[], unsigned int width, unsigned int height) {  // utils.hpp:123
    jb_compat::ck(jb_blockify_planar_i32(jb_compat::ctx(), (const int32_t*)input_arr.data(), width, height, (int32_t*)output_arr));
}
inline void diagonalZigZagBlock(int linear_arr[], int zigzag_arr[]) {          // utils.hpp:126: one block
    jb_compat::ck(jb_zigzag(jb_compat::ctx(), (const int32_t*)linear_arr, (int32_t*)zigzag_arr, 1));
}
inline void access2DArrayRow(int* row, int n) {                                // utils.hpp:124 (declared, never defined)
    for (int i = 0; i < n; ++i) std::cout << row[i] << (i + 1 < n ? " " : "\n");
}
inline void seperateChannels(int zigzag_arr[][64], int y[][64], int cb[][64], int cr[][64], int numRowsPerChannel) {  // utils.hpp:129
    const size_t n = (size_t)numRowsPerChannel * 64 * sizeof(int);     // the array is planar by channel: three copies
    memcpy(y, zigzag_arr, n);
    memcpy(cb, zigzag_arr + numRowsPerChannel, n);
    memcpy(cr, zigzag_arr + 2 * (size_t)numRowsPerChannel, n);
}
// utils.hpp:134-135 on the device function the entropy coder itself uses (jb_value_categories)
inline const int16_t getValueCategory(const int16_t value) {
    uint8_t cat = 0;
    uint16_t bits = 0;
    jb_compat::ck(jb_value_categories(jb_compat::ctx(), &value, 1, &cat, &bits));
    return cat;
}
inline const std::string valueToBitString(const int16_t value) {
    uint8_t cat = 0;
    uint16_t bits = 0;
    jb_compat::ck(jb_value_categories(jb_compat::ctx(), &value, 1, &cat, &bits));
    std::string s(cat, '0');
    for (int i = 0; i < cat; ++i)
        if (bits & (1u << (cat - 1 - i))) s[(size_t)i] = '1';
    return s;
}
inline void RLEBlockAC(int zigzag_array[], std::vector<int>& rle_vector) {     // utils.cpp:572 (one block)
    int32_t pairs[128];
    uint32_t count = 0;
    jb_compat::ck(jb_rle(jb_compat::ctx(), (const int32_t*)zigzag_array, 1, jb_compat::flags(), pairs, &count));
    rle_vector.insert(rle_vector.end(), pairs, pairs + count);
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

// ---- debug printers (utils.hpp:108-115), host side: an 8x8 window of an image / planar vector to stdout ------
inline void printMsg(std::string msg) { if (!msg.empty()) std::cout << "## " << msg << " ##" << std::endl; }
namespace jb_compat {
template <class Get>
inline void preview(size_t x0, size_t y0, size_t nx, size_t ny, const std::string& msg, const char* fmt, Get get) {
    printMsg(msg);
    std::cout << "Previewing pixels from (" << x0 << ", " << y0 << ") to (" << x0 + nx - 1 << ", " << y0 + ny - 1 << "):" << std::endl;
    for (size_t y = y0; y < y0 + ny; ++y) {
        for (size_t x = x0; x < x0 + nx; ++x) {
            double a, b, c;
            get(x, y, a, b, c);
            printf(fmt, a, b, c);
        }
        printf("\n");
    }
}
}  // namespace jb_compat
inline void previewImage(ppm_t* img, size_t x0 = 0, size_t y0 = 0, size_t nx = 8, size_t ny = 8, std::string msg = "") {
    jb_compat::preview(x0, y0, nx, ny, msg, "(%3.0f, %3.0f, %3.0f)   ", [&](size_t x, size_t y, double& a, double& b, double& c) {
        rgb_pixel_t p = getPixel(img, x, y); a = p.r; b = p.g; c = p.b; });
}
inline void previewImageD(ppm_d_t* img, size_t x0 = 0, size_t y0 = 0, size_t nx = 8, size_t ny = 8, std::string msg = "") {
    jb_compat::preview(x0, y0, nx, ny, msg, "(%6.2f,%6.2f,%6.2f) ", [&](size_t x, size_t y, double& a, double& b, double& c) {
        const rgb_pixel_d_t& p = img->data[y * img->width + x]; a = p.r; b = p.g; c = p.b; });
}
template <class T>
inline void jb_compat_preview_planar(std::vector<T>& v, unsigned w, unsigned h, size_t x0, size_t y0, size_t nx, size_t ny,
                                     const std::string& msg, const char* fmt) {
    jb_compat::preview(x0, y0, nx, ny, msg, fmt, [&](size_t x, size_t y, double& a, double& b, double& c) {
        size_t i = y * w + x, n = (size_t)w * h; a = (double)v[i]; b = (double)v[i + n]; c = (double)v[i + 2 * n]; });
}
inline void previewImageLinear(std::vector<cl_uint>& v, const unsigned int w, const unsigned int h, size_t x0 = 0, size_t y0 = 0,
                               size_t nx = 8, size_t ny = 8, std::string msg = "") {
    jb_compat_preview_planar(v, w, h, x0, y0, nx, ny, msg, "(%3.0f, %3.0f, %3.0f)   ");
}
inline void previewImageLinearI(std::vector<int>& v, const unsigned int w, const unsigned int h, size_t x0 = 0, size_t y0 = 0,
                                size_t nx = 8, size_t ny = 8, std::string msg = "") {
    jb_compat_preview_planar(v, w, h, x0, y0, nx, ny, msg, "(%3.0f, %3.0f, %3.0f)   ");
}
inline void previewImageLinearD(std::vector<float>& v, const unsigned int w, const unsigned int h, size_t x0 = 0, size_t y0 = 0,
                                size_t nx = 8, size_t ny = 8, std::string msg = "") {
    jb_compat_preview_planar(v, w, h, x0, y0, nx, ny, msg, "(%6.2f, %6.2f, %6.2f)   ");
}
