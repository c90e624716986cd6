// compat_demo -- the reference's CPU driver sequence (JpegEncoderHost,
// src/OpenCLProject_JpegEncoder.cpp:59-225) written against utils_compat.hpp, i.e.
// against the reference's own function names, with every stage running on the B200
// through libjpegb200.so.  It prints the reference's per-stage timing lines
// (cpp:62,78,144,151,162,173,201,217,229,247) and, for the tests, the bit count and an
// FNV-1a-64 digest of the Huffman bit string and of the zigzag array.
//
//   compat_demo in.ppm [--ref-exact]      --ref-exact = reproduce Q1+Q2+Q3 (as written)
#include <chrono>
#include <cinttypes>

#include "utils_compat.hpp"

static double now_us() {
    using namespace std::chrono;
    return duration<double, std::micro>(steady_clock::now().time_since_epoch()).count();
}

static uint64_t fnv1a(const void* p, size_t n) {
    const uint8_t* b = (const uint8_t*)p;
    uint64_t h = 0xcbf29ce484222325ull;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001b3ull; }
    return h;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s in.ppm [--ref-exact]\n", argv[0]); return 2; }
    if (argc > 2 && !strcmp(argv[2], "--ref-exact"))
        jb_compat::flags() = JB_FLAG_REF_INPLACE_DCT | JB_FLAG_REF_TYPO_TABLES | JB_FLAG_REF_ALWAYS_EOB;
    ppm_t img;
    if (readPPMImage(argv[1], &img.width, &img.height, &img.data) == -1) return 1;
    try {
        jb_compat::ctx();  // create the context outside the timed regions
        double t0, total = 0, dt;
        printf("\n### B200 Implementation (staged, reference function surface) ###\n");
#define STAGE(label, code) t0 = now_us(); code; dt = now_us() - t0; total += dt; printf("%s Time B200: %.0f us\n", label, dt);
        STAGE("CSC", performCSC(&img))
        STAGE("CDS", performCDS(&img))
        size_t nW, nH;
        if (img.width % 8 == 0 && img.height % 8 == 0) { nW = img.width; nH = img.height; }
        else getNearest8x8ImageSize(img.width, img.height, &nW, &nH);
        ppm_t img3{nW, nH, (rgb_pixel_t*)malloc(nW * nH * sizeof(rgb_pixel_t))};
        ppm_d_t imgd{nW, nH, (rgb_pixel_d_t*)malloc(nW * nH * sizeof(rgb_pixel_d_t))};
        STAGE("Total Copy", copyToLargerImage(&img, &img3); addReversedPadding(&img3, img.width, img.height);
              copyUIntToDoubleImage(&img3, &imgd))
        STAGE("Level Shifting", substractfromAll(&imgd, 128.0))
        STAGE("DCT", performDCT(&imgd))
        STAGE("Quantization", performQuantization(&imgd, quant_mat_lum, quant_mat_chrom))
        unsigned rpc = (unsigned)(nW * nH / 64), rows = rpc * 3;
        int(*lin)[64] = (int(*)[64])malloc((size_t)rows * 64 * sizeof(int));
        int(*zz)[64] = (int(*)[64])malloc((size_t)rows * 64 * sizeof(int));
        STAGE("ZigZag", everyMCUisnow2DArray(&imgd, lin); performZigZag(lin, zz, (int)rows))
        std::vector<std::vector<int>> rle;
        STAGE("RLE", performRLE(zz, rle, (int)rows))
        std::string bits;
        STAGE("Huffman", bits = HuffmanEncoder(zz, rle, (int)rpc))
        printf("Total Time B200: %.0f us\n", total);
        printf("padded %zux%zu  zigzag_fnv %016" PRIx64 "  nbits %zu  bits_fnv %016" PRIx64 "\n", nW, nH,
               fnv1a(zz, (size_t)rows * 64 * sizeof(int)), bits.size(), fnv1a(bits.data(), bits.size()));
        free(lin); free(zz); free(img3.data); free(imgd.data);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    jb_compat::shutdown();
    free(img.data);
    return 0;
}
