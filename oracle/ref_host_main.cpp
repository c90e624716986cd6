// TEST INFRASTRUCTURE ONLY.  The reference's own CPU driver function, JpegEncoderHost
// (/root/reference/src/OpenCLProject_JpegEncoder.cpp:28-250), compiled UNMODIFIED -- the lines are extracted by
// oracle/Makefile into oracle/_ref/ref_host_extract.inc at build time, nothing of it is stored in this
// repository -- against host/utils_compat.hpp + host/core_compat.hpp and linked with libjpegb200.so: every stage
// function it calls (performCSC ... HuffmanEncoder, removeRedChannel) then runs on the B200.  This is the proof
// behind INTEGRATION.md's "re-links unchanged"; tests/test_gpu_compat.py runs it on fruit.ppm and checks the PPM
// dumps the function writes (to ../data/, relative to the working directory) against the golden digests.
//
//   ref_host_b200 in.ppm [--ref-exact]
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "utils_compat.hpp"
#include "core_compat.hpp"

#include "ref_host_extract.inc"  // int JpegEncoderHost(ppm_t imgCPU, CPUTelemetry *cpu_telemetry = NULL)

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s in.ppm [--ref-exact]\n", argv[0]); return 2; }
    if (argc > 2 && !strcmp(argv[2], "--ref-exact"))
        jb_compat::flags() = JB_FLAG_REF_INPLACE_DCT | JB_FLAG_REF_TYPO_TABLES | JB_FLAG_REF_ALWAYS_EOB;
    ppm_t img;
    if (readPPMImage(argv[1], &img.width, &img.height, &img.data) == -1) return 1;
    CPUTelemetry t{};
    int rc = 1;
    try {
        jb_compat::ctx();
        rc = JpegEncoderHost(img, &t);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    printf("telemetry_us CSC %.0f CDS %.0f levelShift %.0f DCT %.0f Quant %.0f TotalCopy %.0f zigZag %.0f RLE %.0f Huffman %.0f\n",
           t.CSCTime, t.CDSTime, t.levelShiftTime, t.DCTTime, t.QuantTime, t.TotalCopyTime, t.zigZagTime, t.RLETime, t.HuffmanTime);
    jb_compat::shutdown();
    return rc;
}
