"""Randomised parity sweep through the C ABI on one long-lived context: random sizes, modes, qualities,
restart intervals, batch sizes, contents (smooth, full-range noise, grey noise) and mode flags (CUDA-core kernels,
the reference's in-place transform, optimised Huffman tables), coefficients and JFIF bytes
against the oracle.  Calls of very different shapes on the same context exercise the workspace re-use paths
(arena regrowth, cached tables, the sticky entropy-workspace budget); tests/tools/fuzz_gpu.py is the long form."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [7, 8, 21, 22, 23, 24])
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
            inplace = (not flags) and rng.random() < 0.2  # Q1: the reference's in-place transform, fused path
            optimize = rng.random() < 0.2                 # per-call optimal Huffman tables (single frames: per-image oracle)
            quirks = ol.Q1 if inplace else 0
            flags |= (jb.FLAG_REF_INPLACE_DCT if inplace else 0) | (jb.FLAG_OPTIMIZE_HUFFMAN if optimize else 0)
            N = 1 if optimize else int(rng.choice([1, 1, 2, 3]))
            kind = int(rng.integers(0, 3))
            frames = np.stack([ol.synth(1000 * done + f, W, H) if kind == 0 else
                               rng.integers(0, 256, (H, W, 3), dtype=np.uint8) if kind == 1 else
                               np.repeat(rng.integers(0, 256, (H, W, 1), dtype=np.uint8), 3, axis=2) for f in range(N)])
            ql, qc = ol.quality_tables(q)
            p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri, flags=flags)
            tag = f"case {done}: sub {sub} {W}x{H} q{q} ri {ri} flags {flags} kind {kind} N {N}"
            assert np.array_equal(enc.transform(frames[0], p), ol.transform(frames[0], sub, ql, qc, quirks)), tag
            out, offs, sizes = enc.encode_batch(frames, p, out=np.empty(N * (W * H * 12 + 65536), np.uint8))
            for f in range(N):
                got = bytes(out[int(offs[f]): int(offs[f] + sizes[f])])
                want = (ol.encode_jfif_optimized if optimize else ol.encode_jfif)(frames[f], sub, ql, qc, ri, quirks)
                assert got == want, tag + f" frame {f}"
            done += 1
    finally:
        enc.close()


def test_many_small_frames_odd_pitches_and_device_api(jb):
    """Batches of up to 40 small frames (the units of the tcgen05 kernels wrap over MCU rows and frames), rows /
    frames / first byte at odd offsets (every alignment variant of the kernels), host and device-resident calls."""
    enc = jb.Encoder(0)
    rng = np.random.default_rng(3)
    try:
        done = 0
        while done < 40:
            sub = int(rng.choice([ol.SUB_444, ol.SUB_REPL420, ol.SUB_420]))
            m = 16 if sub == ol.SUB_420 else 8
            W, H = int(rng.integers(m, 200)), int(rng.integers(m, 120))
            if (-W) % m > W or (-H) % m > H:
                continue
            N, q = int(rng.integers(1, 40)), int(rng.choice([50, 75, 95]))
            ri = int(rng.choice([0, 0, 2, -(-W // m)]))
            pad, gap, lead = int(rng.choice([0, 0, 1, 4, 8, 16, 5])), int(rng.choice([0, 0, 3, 16, 64])), int(rng.choice([0, 0, 1, 4, 8]))
            pitch = W * 3 + pad
            stride = pitch * H + gap
            buf = np.zeros(lead + stride * N + 64, np.uint8)
            frames = []
            for f in range(N):
                img = ol.synth(77 * done + f, W, H) if rng.random() < 0.6 else rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
                frames.append(img)
                rows = buf[lead + f * stride: lead + f * stride + pitch * H].reshape(H, pitch)
                rows[:, : W * 3] = img.reshape(H, W * 3)
            ql, qc = ol.quality_tables(q)
            p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri)
            cap = N * (W * H * 6 + 65536)
            out, offs, sizes = np.empty(cap, np.uint8), np.zeros(N, np.uint64), np.zeros(N, np.uint64)
            device = rng.random() < 0.5
            if device:
                d_in, d_out, d_meta = enc.device_alloc(buf.nbytes), enc.device_alloc(cap), enc.device_alloc(16 * N + 16)
                try:
                    enc.h2d(d_in, buf)
                    enc.encode_batch_device(d_in + lead, N, W, H, pitch, stride, p, d_out, cap, d_meta, d_meta + 8 * N, d_meta + 16 * N)
                    enc.sync()
                    meta = np.zeros(2 * N + 1, np.uint64)
                    enc.d2h(out, d_out)
                    enc.d2h(meta, d_meta)
                    offs, sizes = meta[:N], meta[N: 2 * N]
                finally:
                    for d in (d_in, d_out, d_meta):
                        enc.device_free(d)
            else:
                enc.encode_batch_ptr(buf.ctypes.data + lead, N, W, H, pitch, stride, p, out.ctypes.data, cap, offs, sizes)
            for f in range(N):
                got = bytes(out[int(offs[f]): int(offs[f]) + int(sizes[f])])
                assert got == ol.encode_jfif(frames[f], sub, ql, qc, ri), \
                    f"case {done} frame {f}: sub {sub} {W}x{H} N {N} q{q} ri {ri} pad {pad} gap {gap} lead {lead} device {device}"
            done += 1
    finally:
        enc.close()


@pytest.mark.parametrize("shape", [(65535, 16), (16, 65535), (65535, 9), (9, 4000), (8191, 1001), (65528, 24)])
def test_extreme_shapes(jb, shape):
    """Maximal SOF0 width / height, single MCU rows and columns, a large odd size: JFIF bytes == the oracle."""
    W, H = shape
    enc = jb.Encoder(0)
    try:
        for sub in (ol.SUB_420, ol.SUB_444, ol.SUB_REPL420):
            m = 16 if sub == ol.SUB_420 else 8
            if (-W) % m > W or (-H) % m > H:
                continue
            img = ol.synth(W * 7 + H, W, H)
            ql, qc = ol.quality_tables(75)
            ri = min(-(-W // m), 65535) if W * H > 100000 else 0
            p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri)
            assert enc.encode_jfif(img, p, cap=W * H * 3 + (1 << 20)) == ol.encode_jfif(img, sub, ql, qc, ri), f"sub {sub}"
    finally:
        enc.close()


def test_staged_functions_random_sizes(jb):
    """The per-stage entry points (one per reference function) on random image sizes, both DCT variants and both
    entropy table variants: every stage equals the oracle's (= the reference's) stage bit for bit."""
    L = ol.oracle()
    enc = jb.Encoder(0)
    rng = np.random.default_rng(11)
    try:
        for case in range(14):
            W, H = int(rng.integers(8, 150)), int(rng.integers(8, 100))
            img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8) if case % 2 else ol.synth(case, W, H)
            inplace = bool(case % 3 == 0)
            quirks = (ol.Q1 if inplace else 0) | (ol.Q2 | ol.Q3 if case % 4 == 0 else 0)
            a, b = img.copy(), img.copy()
            enc.performCSC(a)
            L.orc_csc(b.reshape(-1), W * H)
            assert np.array_equal(a, b), f"CSC {W}x{H}"
            enc.performCDS(a)
            L.orc_cds(b.reshape(-1), W, H)
            assert np.array_equal(a, b), f"CDS {W}x{H}"
            nW, nH = -(-W // 8) * 8, -(-H // 8) * 8
            pa, pb = enc.padMirror(a), np.zeros((nH, nW, 3), np.uint8)
            assert L.orc_pad_mirror(b.reshape(-1), W, H, pb.reshape(-1), nW, nH) == 0
            assert np.array_equal(pa, pb), f"pad {W}x{H}"
            da, db = enc.copyUIntToDoubleImage(pa), np.zeros(pb.size, np.float64)
            L.orc_u8_to_double(pb.reshape(-1), db, pb.size)
            enc.substractfromAll(da, 128.0)
            L.orc_subtract(db, db.size, 128.0)
            enc.performDCT(da, jb.FLAG_REF_INPLACE_DCT if inplace else 0)
            L.orc_dct_image(db, nW, nH, 1 if inplace else 0)
            assert np.array_equal(da.reshape(-1), db), f"DCT {W}x{H} inplace={inplace}"  # binary64, bit for bit
            ql, qc = ol.quality_tables(int(rng.choice([50, 75, 95])))
            enc.performQuantization(da, ql, qc)
            L.orc_quantize_image(db, nW, nH, ql, qc)
            assert np.array_equal(da.reshape(-1), db), f"quant {W}x{H}"
            rpc = nW * nH // 64
            la, lb = enc.everyMCUisnow2DArray(da), np.zeros((3 * rpc, 64), np.int32)
            L.orc_blockify(db, nW, nH, lb)
            assert np.array_equal(la, lb)
            za, zb = enc.performZigZag(la), np.zeros_like(lb)
            L.orc_zigzag(lb, zb, 3 * rpc)
            assert np.array_equal(za, zb)
            rle = enc.performRLE(za, jb.FLAG_REF_ALWAYS_EOB if quirks & ol.Q3 else 0)
            pairs = np.zeros(130, np.int32)
            for r in range(0, 3 * rpc, max(1, rpc // 5)):
                n = L.orc_rle_block(np.ascontiguousarray(zb[r]), pairs, 1 if quirks & ol.Q3 else 0)
                assert np.array_equal(rle[r], pairs[:n]), f"RLE row {r}"
            flags = (jb.FLAG_REF_TYPO_TABLES if quirks & ol.Q2 else 0) | (jb.FLAG_REF_ALWAYS_EOB if quirks & ol.Q3 else 0)
            bits, nbits = enc.HuffmanEncoder(za, rpc, flags)
            packed = np.zeros(zb.size * 4, np.uint8)
            nb = L.orc_huffman_ref(zb, rpc, quirks & (ol.Q2 | ol.Q3), packed, packed.size)
            assert nbits == nb and np.array_equal(bits, packed[: (nb + 7) // 8]), f"Huffman {W}x{H} quirks {quirks}"
    finally:
        enc.close()


def test_device_api_respects_buffer_bounds(jb):
    """Guard bands around the caller's device buffers (input, output, frame tables) stay untouched, and an output
    buffer that is too small is reported, not overrun (compute-sanitizer is not available on the GPU pool)."""
    enc = jb.Encoder(0)
    rng = np.random.default_rng(17)
    G = 4096
    try:
        for case in range(8):
            sub = int(rng.choice([ol.SUB_444, ol.SUB_REPL420, ol.SUB_420]))
            m = 16 if sub == ol.SUB_420 else 8
            W, H, N = int(rng.integers(m, 300)), int(rng.integers(m, 150)), int(rng.integers(1, 6))
            if (-W) % m > W or (-H) % m > H:
                continue
            frames = np.stack([rng.integers(0, 256, (H, W, 3), dtype=np.uint8) for _ in range(N)])
            p = jb.make_params(sub, quality=int(rng.choice([50, 95])), restart_interval=int(rng.choice([0, 2])))
            want = [enc.encode_jfif(f, p, cap=W * H * 12 + 65536) for f in frames]
            need = sum(len(w) for w in want)
            for cap in (need, need - 1):  # exact fit, then one byte short
                n_in = frames.nbytes
                d_in, d_out, d_meta = enc.device_alloc(n_in + 2 * G), enc.device_alloc(need + 2 * G), enc.device_alloc(16 * N + 8 + 2 * G)
                try:
                    guard = np.full(G, 0xA5, np.uint8)
                    for d, n in ((d_in, n_in), (d_out, need), (d_meta, 16 * N + 8)):
                        enc.h2d(d, guard)
                        enc.h2d(d + G + n, guard)
                    enc.h2d(d_in + G, frames)
                    ok = True
                    try:
                        enc.encode_batch_device(d_in + G, N, W, H, W * 3, W * H * 3, p, d_out + G, cap, d_meta + G, d_meta + G + 8 * N, d_meta + G + 16 * N)
                        enc.sync()
                    except jb.JbError as e:
                        ok = False
                        assert e.code == jb.E_NOSPACE and cap < need
                    assert ok == (cap >= need)
                    back = np.zeros(G, np.uint8)
                    for d, n in ((d_in, n_in), (d_out, need), (d_meta, 16 * N + 8)):
                        for off in (0, G + n):
                            enc.d2h(back, d + off)
                            assert np.array_equal(back, guard), f"guard band overwritten (case {case}, cap {cap})"
                    if ok:
                        out = np.zeros(need, np.uint8)
                        enc.d2h(out, d_out + G)
                        assert bytes(out) == b"".join(want)
                finally:
                    for d in (d_in, d_out, d_meta):
                        enc.device_free(d)
    finally:
        enc.close()
