#!/usr/bin/env python
"""Measure how the B200's tcgen05 fp16 MMA (fp32 accumulation in TMEM) rounds, with the operand layout and MMA order
of k_transform_tc (tests/tools/tc_probe.cu).  Operands are fixed-point numbers, the exact result is computed in
integers, so every deviation is the hardware's.  Prints JSON lines; summarised in profiles/r02_tc_numerics.md.

  hi_exact     products are multiples of a quantum Q and all partial sums stay below 2^24 Q: must be EXACT
  lo_phase     an accumulator of up to 2^23 Q_hi receives 4 MMAs of 16 small products (multiples of q0 = Q_hi / 2^11):
               error per MMA step in ulps of the step's largest magnitude
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
L = C.CDLL(os.path.join(HERE, "..", "_build", "libtcprobe.so"))
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
L.tc_probe.argtypes = [u16p, u16p, u16p, np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"), C.c_int, C.c_int]


def probe(A, B0, B1, n0=4, n1=4):
    """A [128][64], B0/B1 [64][64] as float64 holding fp16-representable values; returns D [128][64] float32."""
    a16, b0, b1 = A.astype(np.float16), B0.astype(np.float16), B1.astype(np.float16)
    assert np.array_equal(a16.astype(np.float64), A) and np.array_equal(b0.astype(np.float64), B0) and np.array_equal(b1.astype(np.float64), B1)
    D = np.zeros((128, 64), np.float32)
    rc = L.tc_probe(a16.view(np.uint16), b0.view(np.uint16), b1.view(np.uint16), D, n0, n1)
    assert rc == 0, rc
    return D


def ulp32(x):
    x = np.abs(np.asarray(x, np.float64))
    e = np.floor(np.log2(np.maximum(x, 2.0 ** -126)))
    return 2.0 ** (e - 23)


def main():
    rng = np.random.default_rng(1)
    out = []
    # ---- 1. hi phase: fixed point, partial sums < 2^24 quanta -> exact ---------------------------------------------
    worst = 0.0
    for trial in range(200):
        e = int(rng.integers(-9, 4))
        kind = trial % 4
        A = rng.integers(-128, 128, (128, 64)).astype(np.float64)
        Bi = rng.integers(-1024, 1025, (64, 64)).astype(np.float64)
        if kind == 1:   # extreme magnitudes, same signs: sums reach 2^23 quanta
            A = np.where(rng.random((128, 64)) < 0.5, -128.0, 127.0)
            Bi = np.sign(A[rng.integers(0, 128)])[None, :] * 1024.0 * np.ones((64, 1))
        if kind == 2:   # per-row quanta (each output row n has its own power-of-two scale)
            Bi = Bi * (2.0 ** rng.integers(-3, 3, (64, 1)))
            Bi = np.clip(Bi, -2048, 2048)
        if kind == 3:   # sparse large + many tiny: alignment stress
            Bi = np.where(rng.random((64, 64)) < 0.1, Bi, np.sign(Bi))
        B = Bi * 2.0 ** e
        D = probe(A, B, np.zeros((64, 64)), 4, 0)
        exact = A @ B.T
        worst = max(worst, float(np.max(np.abs(D.astype(np.float64) - exact) / ulp32(np.maximum(np.abs(exact), 2.0 ** e)))))
    out.append({"experiment": "hi_exact", "trials": 200, "max_error_ulps": worst})
    # ---- 1b. fp16 SUBNORMAL operands (multiples of 2^-24 below 2^-14): are they taken exactly? -----------------------
    worst, flushed = 0.0, 0
    for trial in range(100):
        A = rng.integers(-128, 128, (128, 64)).astype(np.float64)
        Bi = rng.integers(-1023, 1024, (64, 64)).astype(np.float64)
        B = Bi * 2.0 ** -24
        D = probe(A, B, np.zeros((64, 64)), 4, 0)
        exact = A @ B.T
        worst = max(worst, float(np.max(np.abs(D.astype(np.float64) - exact)) * 2.0 ** 24))
        flushed += int(np.sum((D == 0) & (exact != 0)))
    out.append({"experiment": "fp16_subnormal_operands", "trials": 100, "max_error_in_units_of_2^-24": worst, "flushed_to_zero": flushed})
    # ---- 2. lo phase: big accumulator + 4 x 16 small products ------------------------------------------------------
    for order in ("hi_then_lo", "lo_then_hi"):
        stats = {"max_err_over_ulp_final": 0.0, "max_err_over_ulp_acc": 0.0, "rz_exact": 0, "rn_exact": 0, "n": 0}
        for trial in range(400):
            e = int(rng.integers(-3, 6))
            kind = trial % 4
            A = rng.integers(-128, 128, (128, 64)).astype(np.float64)
            Hi = rng.integers(-1024, 1025, (64, 64)).astype(np.float64)
            Lo = rng.integers(-1024, 1025, (64, 64)).astype(np.float64)
            if kind == 1:  # accumulator near the top of its range, lo products all of one sign
                A = np.where(rng.random((128, 64)) < 0.5, -128.0, 127.0)
                Hi = np.sign(A[0])[None, :] * 1024.0 * np.ones((64, 1))
                Lo = np.sign(A[0])[None, :] * rng.integers(1, 1025, (64, 64))
            if kind == 2:  # lo products with all low bits set
                Lo = np.where(rng.random((64, 64)) < 0.5, 1023.0, -1023.0)
                A = rng.choice([-127.0, 127.0, 125.0, -125.0], (128, 64))
            if kind == 3:  # small accumulator (cancellation in the hi phase), large lo part
                Hi = Hi * (rng.random((64, 64)) < 0.05)
            Bh, Bl = Hi * 2.0 ** e, Lo * 2.0 ** (e - 11)
            D = probe(A, Bh, Bl, 4, 4) if order == "hi_then_lo" else probe(A, Bl, Bh, 4, 4)
            hi_part, lo_part = A @ Bh.T, A @ Bl.T
            exact = hi_part + lo_part
            err = np.abs(D.astype(np.float64) - exact)
            stats["max_err_over_ulp_final"] = max(stats["max_err_over_ulp_final"], float(np.max(err / ulp32(exact))))
            # ulp of the largest magnitude any step can see (accumulator before / after, bounded by |hi| + |lo| partials)
            big = np.maximum(np.abs(hi_part), np.abs(exact))
            stats["max_err_over_ulp_acc"] = max(stats["max_err_over_ulp_acc"], float(np.max(err / ulp32(big))))
            f32 = exact.astype(np.float32)  # round to nearest even
            rz = np.where(np.abs(f32.astype(np.float64)) > np.abs(exact), np.nextafter(f32, np.float32(0)), f32)
            stats["rn_exact"] += int(np.sum(D == f32))
            stats["rz_exact"] += int(np.sum(D == rz))
            stats["n"] += D.size
        out.append({"experiment": "lo_phase_" + order, **stats})
    # ---- 3. one MMA step in isolation: accumulator (from one product) + 16 products, per-step error -----------------
    worst = 0.0
    for trial in range(300):
        A = np.zeros((128, 64))
        A[:, :16] = rng.integers(-128, 128, (128, 16))
        Hi = np.zeros((64, 64))
        Hi[:, :16] = rng.integers(-1024, 1025, (64, 16)) * 2.0 ** 4
        Lo = np.zeros((64, 64))
        Lo[:, :16] = rng.integers(-1024, 1025, (64, 16)) * 2.0 ** -7
        D = probe(A, Hi, Lo, 1, 1)      # step 1: 16 hi products (exact), step 2: accumulator + 16 lo products
        hi_part, exact = A @ Hi.T, A @ Hi.T + A @ Lo.T
        err = np.abs(D.astype(np.float64) - exact)
        worst = max(worst, float(np.max(err / ulp32(np.maximum(np.abs(hi_part), np.abs(exact))))))
    out.append({"experiment": "single_step_acc_plus_16_products", "trials": 300, "max_error_ulps_of_step_max": worst})
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    sys.exit(main())
