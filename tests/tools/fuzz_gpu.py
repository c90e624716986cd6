"""Randomised parity sweep on the GPU box: random sizes / modes / qualities / restart intervals / batch sizes,
coefficients and JFIF bytes against the oracle.  python tests/tools/fuzz_gpu.py [cases] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
import oracle_lib as ol
jb = g.load()
enc = jb.Encoder(0)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for c in range(cases):
    sub = int(rng.choice([ol.SUB_444, ol.SUB_REPL420, ol.SUB_420]))
    m = 16 if sub == ol.SUB_420 else 8
    W = int(rng.integers(m, 900)); H = int(rng.integers(m, 260))
    if rng.random() < 0.3: W = int(rng.choice([16, 32, 48, 64, 256, 512, 640, 1920 // 2]))
    if rng.random() < 0.3: H = int(rng.choice([16, 32, 64, 90, 128, 136]))
    if (-W) % m > W or (-H) % m > H: continue
    q = int(rng.choice([10, 50, 75, 90, 100]))
    n_mcu = (-(-W // m)) * (-(-H // m))
    ri = int(rng.choice([0, 0, 1, 3, -(-W // m), 1000]))
    fma = jb.FLAG_FMA_DCT if rng.random() < 0.25 else 0
    inplace = (not fma) and rng.random() < 0.2   # Q1: the reference's in-place transform in the fused path
    optimize = rng.random() < 0.2                # per-call optimal Huffman tables (compared per single frame)
    quirks = ol.Q1 if inplace else 0
    if inplace: fma |= jb.FLAG_REF_INPLACE_DCT
    N = 1 if optimize else int(rng.choice([1, 1, 2, 3]))
    kind = rng.integers(0, 3)
    frames = np.stack([(ol.synth(1000 * c + f, W, H) if kind == 0 else
                        rng.integers(0, 256, (H, W, 3), dtype=np.uint8) if kind == 1 else
                        np.repeat(rng.integers(0, 256, (H, W, 1), dtype=np.uint8), 3, axis=2)) for f in range(N)])
    ql, qc = ol.quality_tables(q)
    p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri, flags=fma | (jb.FLAG_OPTIMIZE_HUFFMAN if optimize else 0))
    ok = True
    for f in range(N):
        got = enc.transform(frames[f], p)
        want = ol.transform(frames[f], sub, ql, qc, quirks)
        if not np.array_equal(got, want):
            ok = False; print("COEF MISMATCH", c, sub, W, H, q, ri, fma, kind, int((got != want).sum()))
    out, offs, sizes = enc.encode_batch(np.ascontiguousarray(frames), p, out=np.empty(N * (W * H * 12 + 65536), np.uint8))
    for f in range(N):
        jf = bytes(out[int(offs[f]): int(offs[f]) + int(sizes[f])])
        want = (ol.encode_jfif_optimized if optimize else ol.encode_jfif)(frames[f], sub, ql, qc, ri, quirks)
        if jf != want:
            a, b = np.frombuffer(jf, np.uint8), np.frombuffer(want, np.uint8)
            n = min(len(a), len(b)); d = np.nonzero(a[:n] != b[:n])[0]
            ok = False; print("JFIF MISMATCH", c, f, sub, W, H, q, ri, fma, kind, "len", len(a), len(b), "first diff", int(d[0]) if len(d) else None, "ndiff", len(d))
            single = enc.encode_jfif(frames[f], p, cap=W * H * 12 + 65536)
            print("   single-frame call equal to oracle:", single == want, " equal to batch:", single == jf)
    bad += not ok
print("fuzz done:", cases, "cases,", bad, "bad")
