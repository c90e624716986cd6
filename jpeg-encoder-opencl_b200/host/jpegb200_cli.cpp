// jpegb200_cli -- command-line driver of the fused path: the counterpart of the
// reference's main() (src/OpenCLProject_JpegEncoder.cpp:255-633), which reads
// ../data/fruit.ppm, runs the stages and prints per-stage times and speed-ups but never
// writes a JPEG.  This one reads any binary P6, encodes on the B200 and writes a JFIF file.
//
//   jpegb200_cli in.ppm out.jpg [--quality Q] [--sub 420|444|repl420] [--restart MCUS] [--repeat N] [--optimize 1]
#include <chrono>

#include "utils_compat.hpp"

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s in.ppm out.jpg [--quality Q] [--sub 420|444|repl420] [--restart MCUS] [--repeat N] [--optimize 1]\n",
                argv[0]);
        return 2;
    }
    int quality = 75, restart = 0, repeat = 1, sub = JB_SUB_420, optimize = 0;
    for (int i = 3; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--quality")) quality = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--restart")) restart = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--repeat")) repeat = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--optimize")) optimize = atoi(argv[i + 1]);  // per-image optimal Huffman tables
        else if (!strcmp(argv[i], "--sub"))
            sub = !strcmp(argv[i + 1], "444") ? JB_SUB_444 : !strcmp(argv[i + 1], "repl420") ? JB_SUB_REPL420 : JB_SUB_420;
    }
    ppm_t img;
    if (readPPMImage(argv[1], &img.width, &img.height, &img.data) == -1) return 1;
    jb_ctx* ctx = nullptr;
    if (jb_create(0, &ctx) != JB_OK) { fprintf(stderr, "no CUDA device (there is no CPU fallback)\n"); return 1; }
    jb_params p{};
    p.subsampling = sub;
    p.restart_interval = restart;
    p.flags = optimize ? JB_FLAG_OPTIMIZE_HUFFMAN : 0u;
    jb_quality_tables(quality, p.qlum, p.qchrom);
    size_t cap = img.width * img.height * 3 + 65536, n = 0;
    // pinned host buffers (jb_host_alloc): the copies are then real DMA transfers instead of staged ones
    void *pin_in = nullptr, *pin_out = nullptr;
    if (jb_host_alloc(&pin_in, img.width * img.height * 3) != JB_OK || jb_host_alloc(&pin_out, cap) != JB_OK) {
        fprintf(stderr, "pinned allocation failed\n");
        return 1;
    }
    memcpy(pin_in, img.data, img.width * img.height * 3);
    uint8_t* out = (uint8_t*)pin_out;
    jb_set_profiling(ctx, 1);
    double best = 1e30;
    for (int r = 0; r < repeat; ++r) {
        jb_reset_counters(ctx);
        auto t0 = std::chrono::steady_clock::now();
        int rc = jb_encode_jfif(ctx, (const uint8_t*)pin_in, img.width, img.height, img.width * 3, &p, out, cap, &n);
        double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        if (rc != JB_OK) { fprintf(stderr, "encode failed: %s\n", jb_last_error(ctx)); return 1; }
        if (us < best) best = us;
    }
    jb_timings t;
    jb_get_timings(ctx, &t);
    printf("%zux%zu -> %zu bytes (%.3f bits/px)\n", img.width, img.height, n, 8.0 * n / (img.width * img.height));
    printf("Transform (CSC+CDS+shift+DCT+quant+zigzag) Time B200: %.1f us\n", t.transform_us);
    printf("Tie fix-up Time B200: %.1f us (%llu coefficients)\n", t.fixup_us, (unsigned long long)t.tie_fixups);
    printf("RLE+Huffman+packing Time B200: %.1f us\n", t.entropy_us);
    printf("Total Copy Time B200: %.1f us\n", t.TotalCopyTime);
    printf("End-to-end (host to host): %.1f us = %.2f MP/s\n", best, img.width * img.height / best);
    FILE* fp = fopen(argv[2], "wb");
    if (!fp) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    fwrite(out, 1, n, fp);
    fclose(fp);
    jb_host_free(pin_in);
    jb_host_free(pin_out);
    jb_destroy(ctx);
    free(img.data);
    return 0;
}
