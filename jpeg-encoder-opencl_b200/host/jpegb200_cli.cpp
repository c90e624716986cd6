// jpegb200_cli -- command-line driver: the counterpart of the reference's main()
// (src/OpenCLProject_JpegEncoder.cpp:255-633), which reads ../data/fruit.ppm, runs the CPU stages, then the GPU
// stages one by one, prints "<stage> time (GPU)" lines and a "## Speedups: ##" table (cpp:621-629) -- but stops at RLE
// (no Huffman on the GPU), leaves the transfers out of the table and never writes a JPEG.  This one reads any binary
// P6, encodes on the B200, writes a JFIF file and prints the same kind of report with Huffman and transfers included.
//
//   jpegb200_cli in.ppm out.jpg [--quality Q] [--sub 420|444|repl420] [--restart MCUS] [--repeat N] [--optimize 1]
//                               [--staged 1] [--cpu-telemetry FILE]
//
//   default         the fused path (jb_encode_jfif): one transform kernel + the entropy coder
//   --staged 1      additionally runs the reference's stage sequence (JpegEncoderHost, cpp:59-225) through the staged
//                   entry points, one GPU kernel per reference function, and reports every stage's kernel time like the
//                   reference does; the staged zigzag array must equal the fused path's coefficients (checked)
//   --cpu-telemetry FILE   per-stage CPU microseconds of the reference's CPU path ("CSCTime 301.5" ... one CPUTelemetry
//                   field per line, e.g. written from oracle/_ref by the tests): enables the speed-up table.  The tool itself
//                   has no CPU implementation to time -- there is no CPU path in this library.
#include <chrono>
#include <map>
#include <string>

#include "utils_compat.hpp"

static double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static bool read_cpu_telemetry(const char* path, CPUTelemetry* t) {
    FILE* fp = fopen(path, "r");
    if (!fp) return false;
    std::map<std::string, double*> f = {{"CSCTime", &t->CSCTime}, {"CDSTime", &t->CDSTime}, {"levelShiftTime", &t->levelShiftTime},
                                         {"DCTTime", &t->DCTTime}, {"QuantTime", &t->QuantTime}, {"TotalCopyTime", &t->TotalCopyTime},
                                         {"zigZagTime", &t->zigZagTime}, {"RLETime", &t->RLETime}, {"HuffmanTime", &t->HuffmanTime}};
    char name[64];
    double v;
    int n = 0;
    while (fscanf(fp, "%63s %lf", name, &v) == 2)
        if (f.count(name)) { *f[name] = v; ++n; }
    fclose(fp);
    return n == 9;
}

static void speedup(const char* label, double cpu_us, double gpu_us) {
    if (gpu_us > 0) printf("%s: %.1f\n", label, cpu_us / gpu_us);
    else printf("%s: n/a\n", label);
}

int main(int argc, char** argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s in.ppm out.jpg [--quality Q] [--sub 420|444|repl420] [--restart MCUS] [--repeat N] [--optimize 1] "
                        "[--staged 1] [--cpu-telemetry FILE]\n", argv[0]);
        return 2;
    }
    int quality = 75, restart = 0, repeat = 1, sub = JB_SUB_420, optimize = 0, staged = 0;
    const char* cpu_file = nullptr;
    for (int i = 3; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--quality")) quality = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--restart")) restart = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--repeat")) repeat = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--optimize")) optimize = atoi(argv[i + 1]);  // per-image optimal Huffman tables
        else if (!strcmp(argv[i], "--staged")) staged = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--cpu-telemetry")) cpu_file = argv[i + 1];
        else if (!strcmp(argv[i], "--sub"))
            sub = !strcmp(argv[i + 1], "444") ? JB_SUB_444 : !strcmp(argv[i + 1], "repl420") ? JB_SUB_REPL420 : JB_SUB_420;
    }
    ppm_t img;
    if (readPPMImage(argv[1], &img.width, &img.height, &img.data) == -1) return 1;
    CPUTelemetry cpu{};
    const bool have_cpu = cpu_file && read_cpu_telemetry(cpu_file, &cpu);
    if (cpu_file && !have_cpu) fprintf(stderr, "warning: %s does not hold the nine CPUTelemetry fields; no speed-up table\n", cpu_file);
    jb_ctx* ctx = nullptr;
    try {
        ctx = jb_compat::ctx();  // raises without a CUDA device: there is no CPU fallback
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    jb_params p{};
    p.subsampling = sub;
    p.restart_interval = restart;
    p.flags = optimize ? JB_FLAG_OPTIMIZE_HUFFMAN : 0u;
    jb_quality_tables(quality, p.qlum, p.qchrom);
    const size_t npx = img.width * img.height;
    size_t cap = npx * 3 + 65536, n = 0;
    // pinned host buffers (jb_host_alloc): the copies are then real DMA transfers instead of staged ones
    void *pin_in = nullptr, *pin_out = nullptr;
    if (jb_host_alloc(&pin_in, npx * 3) != JB_OK || jb_host_alloc(&pin_out, cap) != JB_OK) {
        fprintf(stderr, "pinned allocation failed\n");
        return 1;
    }
    memcpy(pin_in, img.data, npx * 3);
    uint8_t* out = (uint8_t*)pin_out;
    jb_set_profiling(ctx, 1);

    // ---- fused path ---------------------------------------------------------------------------------------------
    double best = 1e30;
    for (int r = 0; r < repeat; ++r) {
        jb_reset_counters(ctx);
        double t0 = now_us();
        int rc = jb_encode_jfif(ctx, (const uint8_t*)pin_in, img.width, img.height, img.width * 3, &p, out, cap, &n);
        double us = now_us() - t0;
        if (rc != JB_OK) { fprintf(stderr, "encode failed: %s\n", jb_last_error(ctx)); return 1; }
        if (us < best) best = us;
    }
    jb_timings t;
    jb_get_timings(ctx, &t);
    printf("%zux%zu -> %zu bytes (%.3f bits/px)\n", img.width, img.height, n, 8.0 * n / npx);
    printf("\n### B200 Implementation (fused path) ###\n");
    printf("Transform (CSC+CDS+shift+DCT+quant+zigzag) Time B200: %.1f us\n", t.transform_us + t.edge_us);
    printf("Tie fix-up Time B200: %.1f us (%llu coefficients)\n", t.fixup_us, (unsigned long long)t.tie_fixups);
    printf("RLE+Huffman+packing Time B200: %.1f us\n", t.entropy_us);
    printf("Host to device Time B200: %.1f us, device to host: %.1f us\n", t.h2d_us, t.d2h_us);
    printf("End-to-end (host to host): %.1f us = %.2f MP/s\n", best, npx / best);
    if (have_cpu) {
        const double cpu_transform = cpu.CSCTime + cpu.CDSTime + cpu.TotalCopyTime + cpu.levelShiftTime + cpu.DCTTime + cpu.QuantTime + cpu.zigZagTime;
        const double cpu_total = cpu_transform + cpu.RLETime + cpu.HuffmanTime;
        printf("\n## Speedups (fused path): ##\n");
        speedup("CSC..ZigZag (one fused kernel + tie fix-up)", cpu_transform, t.transform_us + t.edge_us + t.fixup_us);
        speedup("RLE + Huffman (entropy coder incl. byte packing, stuffing, markers)", cpu.RLETime + cpu.HuffmanTime, t.entropy_us);
        speedup("All kernels", cpu_total, t.transform_us + t.edge_us + t.fixup_us + t.entropy_us);
        speedup("End to end incl. transfers and launch overheads", cpu_total, best);
    }
    FILE* fp = fopen(argv[2], "wb");
    if (!fp) { fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    fwrite(out, 1, n, fp);
    fclose(fp);

    // ---- the reference's stage sequence, one GPU kernel per reference function -----------------------------------------
    int rc_staged = 0;
    if (staged) {
        try {
            jb_reset_counters(ctx);
            const double w0 = now_us();
            ppm_t work{img.width, img.height, (rgb_pixel_t*)malloc(npx * 3)};
            memcpy(work.data, img.data, npx * 3);
            performCSC(&work);                                              // cpp:59
            performCDS(&work);                                              // cpp:75
            size_t nW = work.width, nH = work.height;
            if (nW % 8 || nH % 8) getNearest8x8ImageSize(work.width, work.height, &nW, &nH);  // cpp:93-98
            ppm_t padded{nW, nH, (rgb_pixel_t*)malloc(nW * nH * sizeof(rgb_pixel_t))};
            ppm_d_t imgd{nW, nH, (rgb_pixel_d_t*)malloc(nW * nH * sizeof(rgb_pixel_d_t))};
            copyToLargerImage(&work, &padded);                              // cpp:109-120
            copyUIntToDoubleImage(&padded, &imgd);                          // cpp:139
            substractfromAll(&imgd, 128.0);                                 // cpp:147
            performDCT(&imgd);                                              // cpp:158 (true DCT; see jb_compat::flags())
            performQuantization(&imgd, (const unsigned int(*)[8])p.qlum, (const unsigned int(*)[8])p.qchrom);  // cpp:169
            const unsigned rpc = (unsigned)(nW * nH / 64), rows = rpc * 3;
            int(*lin)[64] = (int(*)[64])malloc((size_t)rows * 64 * sizeof(int));
            int(*zz)[64] = (int(*)[64])malloc((size_t)rows * 64 * sizeof(int));
            everyMCUisnow2DArray(&imgd, lin);                               // cpp:194
            performZigZag(lin, zz, (int)rows);                              // cpp:197
            std::vector<std::vector<int>> rle;
            performRLE(zz, rle, (int)rows);                                 // cpp:213
            std::string bits = HuffmanEncoder(zz, rle, (int)rpc);           // cpp:225
            const double wall = now_us() - w0;
            jb_timings s;
            jb_get_timings(ctx, &s);
            printf("\n### B200 Implementation (staged: the reference's stage sequence, kernel time per stage) ###\n");
            printf("Color conversion time (GPU): %.1f us\n", s.CSCTime);
            printf("Chroma subsampling time (GPU): %.1f us\n", s.CDSTime);
            printf("Copy + padding + conversion time (GPU): %.1f us\n", s.staged_copy_us);
            printf("Level shifting time (GPU): %.1f us\n", s.levelShiftTime);
            printf("DCT time (GPU): %.1f us\n", s.staged_dct_us);
            printf("Quantization time (GPU): %.1f us\n", s.QuantTime);
            printf("ZigZag time (GPU): %.1f us\n", s.zigZagTime);
            printf("RLE time (GPU): %.1f us\n", s.RLETime);
            printf("Huffman time (GPU): %.1f us   (%zu bits; the reference has no GPU Huffman stage, cpp:623-629)\n", s.entropy_us, bits.size());
            printf("Transfers (per-stage host<->device round trips, as in the reference's main): h2d %.1f us, d2h %.1f us\n", s.h2d_us, s.d2h_us);
            printf("Staged sequence wall time: %.1f us\n", wall);
            if (have_cpu) {
                printf("\n## Speedups: ##\n");
                speedup("Color conversion", cpu.CSCTime, s.CSCTime);
                speedup("Chroma subsampling", cpu.CDSTime, s.CDSTime);
                speedup("Copy + padding", cpu.TotalCopyTime, s.staged_copy_us);
                speedup("Level shifting", cpu.levelShiftTime, s.levelShiftTime);
                speedup("DCT", cpu.DCTTime, s.staged_dct_us);
                speedup("Quantization", cpu.QuantTime, s.QuantTime);
                speedup("ZigZag", cpu.zigZagTime, s.zigZagTime);
                speedup("RLE", cpu.RLETime, s.RLETime);
                speedup("Huffman", cpu.HuffmanTime, s.entropy_us);
                const double cpu_total = cpu.CSCTime + cpu.CDSTime + cpu.TotalCopyTime + cpu.levelShiftTime + cpu.DCTTime + cpu.QuantTime +
                                         cpu.zigZagTime + cpu.RLETime + cpu.HuffmanTime;
                speedup("Whole pipeline incl. transfers (wall)", cpu_total, wall);
            }
            // the staged zigzag array must be the fused path's coefficients (replicated 4:2:0 = the reference's own mode)
            jb_params pr = p;
            pr.subsampling = JB_SUB_REPL420;
            pr.flags = 0;
            pr.restart_interval = 0;
            std::vector<int16_t> coef((size_t)rows * 64);
            jb_compat::ck(jb_transform(ctx, (const uint8_t*)pin_in, img.width, img.height, img.width * 3, &pr, coef.data()));
            size_t bad = 0;
            for (unsigned m = 0; m < rpc; ++m)
                for (int c = 0; c < 3; ++c)
                    for (int k = 0; k < 64; ++k) bad += coef[((size_t)m * 3 + c) * 64 + k] != (int16_t)zz[m + (size_t)rpc * c][k];
            printf("staged zigzag array == fused coefficients: %s (%zu differences)\n", bad ? "NO" : "yes", bad);
            rc_staged = bad ? 1 : 0;
            free(lin); free(zz); free(work.data); free(padded.data); free(imgd.data);
        } catch (const std::exception& e) {
            fprintf(stderr, "staged run failed: %s\n", e.what());
            rc_staged = 1;
        }
    }
    jb_host_free(pin_in);
    jb_host_free(pin_out);
    jb_compat::shutdown();
    free(img.data);
    return rc_staged;
}
