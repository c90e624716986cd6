"""Randomised parity sweep through the C ABI on one long-lived context: random sizes, modes, qualities,
restart intervals, batch sizes and contents (smooth, full-range noise, grey noise), coefficients and JFIF bytes
against the oracle.  Calls of very different shapes on the same context exercise the workspace re-use paths
(arena regrowth, cached tables, the sticky entropy-workspace budget); experiments/fuzz_gpu.py is the long form."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [7, 8])
def test_random_shapes_modes_and_contents(jb, seed):
    enc = jb.Encoder(0)
    rng = np.random.default_rng(seed)
    try:
        done = 0
        while done < 45:
            sub = int(rng.choice([ol.SUB_444, ol.SUB_REPL420, ol.SUB_420]))
            m = 16 if sub == ol.SUB_420 else 8
            W, H = int(rng.integers(m, 700)), int(rng.integers(m, 200))
            if rng.random() < 0.3:
                W = int(rng.choice([16, 32, 48, 64, 256, 512, 640]))
            if (-W) % m > W or (-H) % m > H:
                continue
            q = int(rng.choice([10, 50, 75, 90, 100]))
            ri = int(rng.choice([0, 0, 1, 3, -(-W // m), 1000]))
            flags = jb.FLAG_FMA_DCT if rng.random() < 0.25 else 0
            N = int(rng.choice([1, 1, 2, 3]))
            kind = int(rng.integers(0, 3))
            frames = np.stack([ol.synth(1000 * done + f, W, H) if kind == 0 else
                               rng.integers(0, 256, (H, W, 3), dtype=np.uint8) if kind == 1 else
                               np.repeat(rng.integers(0, 256, (H, W, 1), dtype=np.uint8), 3, axis=2) for f in range(N)])
            ql, qc = ol.quality_tables(q)
            p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri, flags=flags)
            tag = f"case {done}: sub {sub} {W}x{H} q{q} ri {ri} flags {flags} kind {kind} N {N}"
            assert np.array_equal(enc.transform(frames[0], p), ol.transform(frames[0], sub, ql, qc)), tag
            out, offs, sizes = enc.encode_batch(frames, p, out=np.empty(N * (W * H * 12 + 65536), np.uint8))
            for f in range(N):
                got = bytes(out[int(offs[f]): int(offs[f] + sizes[f])])
                assert got == ol.encode_jfif(frames[f], sub, ql, qc, ri), tag + f" frame {f}"
            done += 1
    finally:
        enc.close()
