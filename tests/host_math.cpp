// TEST-ONLY host harness (compiled by tests/test_math_host.py with g++, no GPU needed).
// jb_math.h and jb_tables.cpp are written so that the host evaluates bit-for-bit the
// arithmetic the kernels evaluate; this program checks that arithmetic against the
// reference's formulas (restated here in binary64, utils.cpp:100-109, 314-347, 454-467):
//   1. colour conversion over all 2^24 colours (fixed point + tie table == binary64 + truncation)
//   2. the analytic error bound of the binary32 AAN transform against measured errors
//   3. the block pipeline (DCT -> quantise with near-tie flags): every coefficient the fast
//      path does NOT flag equals the binary64 result; flagged ones are the only candidates
//   4. the integer DC rule for every quantiser value
// It prints one JSON object.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "jb_internal.h"

using namespace jb;

static double cs[8][8];

static void ref_dct_quant(const int* smp, const uint32_t* q, int* out /* natural order */) {
    for (size_t u = 0; u < 8; ++u)
        for (size_t v = 0; v < 8; ++v) {
            double au = (u == 0) ? 1.0 / std::sqrt(2) : 1.0, av = (v == 0) ? 1.0 / std::sqrt(2) : 1.0;
            double s = 0.0;
            for (size_t y = 0; y < 8; ++y)
                for (size_t x = 0; x < 8; ++x) s += (double)smp[y * 8 + x] * cs[u][x] * cs[v][y];
            s *= (au * av / 4.0);
            out[v * 8 + u] = (int)std::round(s / q[v * 8 + u]);
        }
}

// ---- 5. the tcgen05 transform under the measured datapath model (jb_tables.cpp: build_tc_matrices) --------------
// fp16 bit pattern -> double (normal and subnormal)
static double h2d(uint16_t h) {
    int s = h >> 15, e = (h >> 10) & 31, f = h & 1023;
    double v = e ? std::ldexp(1.0 + f / 1024.0, e - 15) : std::ldexp((double)f, -24);
    return s ? -v : v;
}
// one MMA step: accumulator + 16 exact products, aligned to the largest exponent with `guard` guard bits, every
// addend truncated toward zero, exact sum, result truncated toward zero to 24 significant bits
static double mma_step(double acc, const double* prod, int guard) {
    double m = std::fabs(acc);
    for (int i = 0; i < 16; ++i) m = std::fmax(m, std::fabs(prod[i]));
    if (m == 0) return 0;
    const double g = std::ldexp(1.0, std::ilogb(m) - 23 - guard);
    double s = std::trunc(acc / g);
    for (int i = 0; i < 16; ++i) s += std::trunc(prod[i] / g);
    s *= g;
    if (s == 0) return 0;
    const double u = std::ldexp(1.0, std::ilogb(s) - 23);
    return std::trunc(s / u) * u;
}
struct TcReport {
    long coefs = 0, flagged = 0, unflagged_wrong = 0;
    double worst_err_over_bound = 0, min_band = 1;
};
static void tc_model_check(TcReport& rep) {
    const int qualities[5] = {10, 50, 75, 90, 100};
    std::vector<uint8_t> mat(32768);
    srand(777);
    for (int qi = 0; qi < 5; ++qi) {
        uint32_t ql[64], qc[64];
        jb_quality_tables(qualities[qi], ql, qc);
        float tband[2][64];
        build_tc_matrices(ql, qc, JB_TC_STEP_ULPS, 0, 0, mat.data(), tband);
        for (int t = 0; t < 2; ++t) {
            const uint32_t* q = t ? qc : ql;
            double hi[64][64], lo[64][64];
            for (int n = 0; n < 64; ++n)
                for (int k = 0; k < 64; ++k) {
                    size_t off = (size_t)n * 128 + (size_t)(((k >> 3) ^ (n & 7)) << 4) + (size_t)(k & 7) * 2;
                    uint16_t a, b;
                    memcpy(&a, mat.data() + ((size_t)t * 2 + 0) * 8192 + off, 2);
                    memcpy(&b, mat.data() + ((size_t)t * 2 + 1) * 8192 + off, 2);
                    hi[n][k] = h2d(a);
                    lo[n][k] = h2d(b);
                }
            for (int n = 0; n < 64; ++n) rep.min_band = std::fmin(rep.min_band, (double)tband[t][n]);
            for (int trial = 0; trial < 1500; ++trial) {
                int smp[64];
                const int mode = trial % 5;
                for (int i = 0; i < 64; ++i)
                    smp[i] = mode == 0 ? rand() % 256 - 128 : mode == 1 ? ((rand() & 1) ? 127 : -128)
                             : mode == 2 ? (rand() % 9 - 4) + (i % 8) * 16 - 60 : mode == 3 ? (trial / 5) % 256 - 128
                                                                                            : ((rand() & 3) ? -128 : 127);
                int want[64];
                ref_dct_quant(smp, q, want);
                for (int n = 1; n < 64; ++n) {  // (the DC coefficient takes the integer path)
                    double acc = 0, prod[16];
                    for (int phase = 0; phase < 2; ++phase)
                        for (int c = 0; c < 4; ++c) {
                            for (int i = 0; i < 16; ++i) prod[i] = smp[c * 16 + i] * (phase ? lo[n][c * 16 + i] : hi[n][c * 16 + i]);
                            acc = mma_step(acc, prod, 3);
                        }
                    // exact scaled quotient, for the error ratio
                    const int nat = kZigzag[n], v = nat >> 3, u = nat & 7;
                    double s = 0;
                    for (int y = 0; y < 8; ++y)
                        for (int x = 0; x < 8; ++x) s += smp[y * 8 + x] * cs[u][x] * cs[v][y];
                    s *= ((u == 0 ? 1.0 / std::sqrt(2) : 1.0) * (v == 0 ? 1.0 / std::sqrt(2) : 1.0) / 4.0) / q[nat] * JB_TC_W_SCALE;
                    const double bound = (0.5 - 1e-6 - tband[t][n]) * JB_TC_W_SCALE;
                    rep.worst_err_over_bound = std::fmax(rep.worst_err_over_bound, std::fabs(acc - s) / bound);
                    // the epilogue of the kernel (tc_quant_stage): round, distance from the integer, band test
                    const float x = (float)acc, inv = (float)(1.0 / JB_TC_W_SCALE);
                    const float r = fmaf(x, inv, JB_ROUND_MAGIC), d = fmaf(x, inv, JB_ROUND_MAGIC - r);
                    const bool tie = std::fabs(d) > tband[t][n];
                    const int got = (int)(r - JB_ROUND_MAGIC);
                    ++rep.coefs;
                    if (tie) ++rep.flagged;
                    else if (got != want[nat]) ++rep.unflagged_wrong;
                }
            }
        }
    }
}

int main() {
    for (size_t u = 0; u < 8; ++u)
        for (size_t x = 0; x < 8; ++x) cs[u][x] = std::cos((2 * x + 1) * u * M_PI / 16.0);

    // ---- 1. colour conversion ---------------------------------------------------------
    std::vector<uint32_t> ydown(2048);
    build_ydown(ydown.data());
    long csc_bad = 0, y_ties = 0, y_down = 0;
    for (uint32_t r = 0; r < 256; ++r)
        for (uint32_t g = 0; g < 256; ++g)
            for (uint32_t b = 0; b < 256; ++b) {
                uint8_t y = (uint8_t)(0.299 * r + 0.587 * g + 0.114 * b);
                uint8_t cb = (uint8_t)(-0.168736 * r - 0.331264 * g + 0.5 * b + 128);
                uint8_t cr = (uint8_t)(0.5 * r - 0.418688 * g - 0.081312 * b + 128);
                if (csc_y(r, g, b, ydown.data()) != y || csc_cb(r, g, b) != cb || csc_cr(r, g, b) != cr) ++csc_bad;
                if ((csc_ty(r, g, b) & Y_TIE_MASK) == 0) {
                    ++y_ties;
                    if ((csc_ty(r, g, b) >> 24) != y) ++y_down;
                }
            }

    // ---- 2./3. transform error and near-tie logic ---------------------------------------
    double err[64], amax[64];
    aan_error_bound(err, amax);
    double worst_ratio = 0, max_err = 0, max_bound = 0;
    long coefs = 0, flagged = 0, unflagged_wrong = 0, flagged_differ = 0, max_lsb = 0;
    srand(12345);
    const int qualities[4] = {50, 75, 90, 100};
    for (int trial = 0; trial < 60000; ++trial) {
        int smp[64];
        int mode = trial % 6;
        for (int i = 0; i < 64; ++i) {
            int v;
            if (mode == 0) v = rand() % 256 - 128;
            else if (mode == 1) v = (rand() & 1) ? 127 : -128;
            else if (mode == 2) v = ((i * 7 + trial) % 3 == 0) ? 127 : -128;
            else if (mode == 3) v = (rand() % 9 - 4) + (i % 8) * 16 - 60;
            else if (mode == 4) v = (trial / 6) % 256 - 128;  // flat blocks: exact DC ties
            else v = (rand() & 3) ? -128 : 127;
            smp[i] = v;
        }
        float a[64];
        for (int i = 0; i < 64; ++i) a[i] = (float)smp[i];
        for (int r = 0; r < 8; ++r)
            fdct8(a[r * 8], a[r * 8 + 1], a[r * 8 + 2], a[r * 8 + 3], a[r * 8 + 4], a[r * 8 + 5], a[r * 8 + 6], a[r * 8 + 7]);
        for (int c = 0; c < 8; ++c) fdct8(a[c], a[8 + c], a[16 + c], a[24 + c], a[32 + c], a[40 + c], a[48 + c], a[56 + c]);
        for (int v = 0; v < 8; ++v)
            for (int u = 0; u < 8; ++u) {
                double s = 0;
                for (int y = 0; y < 8; ++y)
                    for (int x = 0; x < 8; ++x) s += smp[y * 8 + x] * cs[u][x] * cs[v][y];
                double e = std::fabs((double)a[v * 8 + u] - s * aan_scale(u) * aan_scale(v));
                if (e > max_err) max_err = e;
                if (err[v * 8 + u] > 0 && e / err[v * 8 + u] > worst_ratio) worst_ratio = e / err[v * 8 + u];
                // outputs that only add integers must be exact (up to the binary64 noise of this check)
                if (err[v * 8 + u] == 0 && e > 1e-9) worst_ratio = 1e9;
            }
        // quantise with the tables of one quality (IJG scaling of the reference's q50 tables)
        uint32_t ql[64], qc[64];
        jb_quality_tables(qualities[trial % 4], ql, qc);
        QuantConst k;
        build_quant_const(ql, qc, &k);
        int want[64];
        ref_dct_quant(smp, ql, want);
        for (int n = 0; n < 64; ++n) {
            bool tie = false;
            int got;
            if (n == 0 && k.dc_exact) {
                int S = (int)lrintf(a[0]);
                uint32_t A = (uint32_t)std::abs(S);
                uint32_t m = (uint32_t)(((uint64_t)(2 * A + k.dc_d[0] - 1) * k.dc_m[0]) >> 32);
                got = S < 0 ? -(int)m : (int)m;
            } else {
                uint32_t bits = quantize_bits(a[n], k.mul[0][n], k.band[0][n], tie);
                got = (int)(int16_t)(bits & 0xFFFF);
            }
            ++coefs;
            if (tie) {
                ++flagged;
                if (got != want[n]) ++flagged_differ;
            } else if (got != want[n]) {
                ++unflagged_wrong;
            }
            long d = std::labs((long)got - want[n]);
            if (d > max_lsb) max_lsb = d;
        }
    }
    for (int n = 0; n < 64; ++n)
        if (err[n] > max_bound) max_bound = err[n];

    // ---- 4. integer DC rule for every quantiser value ------------------------------------
    int dc_rule_failures = 0;
    for (uint32_t q = 1; q <= 255; ++q) {
        uint32_t ql[64], qc[64];
        for (int i = 0; i < 64; ++i) ql[i] = qc[i] = q;
        QuantConst k;
        build_quant_const(ql, qc, &k);
        if (!k.dc_exact) ++dc_rule_failures;
    }

    TcReport tc;
    tc_model_check(tc);
    // the band of every matrix variant the library builds (true / in-place transform, 64 samples / 16 cell means per
    // chroma block): a row whose band collapses is replayed for every block (round 2 shipped one such defect)
    double min_band_modes = 1;
    {
        std::vector<uint8_t> mat(32768);
        const int qualities[5] = {10, 50, 75, 90, 100};
        for (int qi = 0; qi < 5; ++qi)
            for (int repl = 0; repl < 2; ++repl)
                for (int inplace = 0; inplace < 2; ++inplace) {
                    uint32_t ql[64], qc[64];
                    jb_quality_tables(qualities[qi], ql, qc);
                    float tband[2][64];
                    build_tc_matrices(ql, qc, JB_TC_STEP_ULPS, repl, inplace, mat.data(), tband);
                    for (int t = 0; t < 2; ++t)
                        for (int n = 1; n < 64; ++n) min_band_modes = std::fmin(min_band_modes, (double)tband[t][n]);
                }
    }

    printf("{\"csc_mismatches\": %ld, \"y_ties\": %ld, \"y_ties_down\": %ld, \"max_err\": %.6e, \"max_bound\": %.6e, "
           "\"worst_err_over_bound\": %.4f, \"coefs\": %ld, \"flagged\": %ld, \"unflagged_wrong\": %ld, "
           "\"flagged_differ\": %ld, \"max_lsb\": %ld, \"dc_rule_failures\": %d, "
           "\"tc_coefs\": %ld, \"tc_flagged\": %ld, \"tc_unflagged_wrong\": %ld, \"tc_worst_err_over_bound\": %.4f, \"tc_min_band\": %.6f, \"tc_min_band_all_modes\": %.6f}\n",
           csc_bad, y_ties, y_down, max_err, max_bound, worst_ratio, coefs, flagged, unflagged_wrong, flagged_differ,
           max_lsb, dc_rule_failures, tc.coefs, tc.flagged, tc.unflagged_wrong, tc.worst_err_over_bound, tc.min_band, min_band_modes);
    return 0;
}
