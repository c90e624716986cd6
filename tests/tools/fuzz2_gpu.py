"""Second randomised sweep: many small frames per batch (units of the tcgen05 kernels wrap over MCU rows and
frames), odd pitches / frame strides / base offsets (alignment variants of the kernels), device-resident API."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import __graft_entry__ as g
import oracle_lib as ol
jb = g.load()
enc = jb.Encoder(0)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for c in range(cases):
    sub = int(rng.choice([ol.SUB_444, ol.SUB_REPL420, ol.SUB_420]))
    m = 16 if sub == ol.SUB_420 else 8
    W = int(rng.integers(m, 200)); H = int(rng.integers(m, 120))
    if (-W) % m > W or (-H) % m > H: continue
    N = int(rng.integers(1, 40))
    q = int(rng.choice([50, 75, 95]))
    ri = int(rng.choice([0, 0, 2, -(-W // m)]))
    pad = int(rng.choice([0, 0, 1, 4, 8, 16, 5]))          # extra bytes per row
    gap = int(rng.choice([0, 0, 3, 16, 64]))               # extra bytes between frames
    lead = int(rng.choice([0, 0, 1, 4, 8]))                # offset of the first byte
    pitch = W * 3 + pad
    stride = pitch * H + gap
    buf = np.zeros(lead + stride * N + 64, np.uint8)
    frames = []
    for f in range(N):
        img = ol.synth(77 * c + f, W, H) if rng.random() < 0.6 else rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        frames.append(img)
        for y in range(H):
            o = lead + f * stride + y * pitch
            buf[o: o + W * 3] = img[y].reshape(-1)
    ql, qc = ol.quality_tables(q)
    p = jb.make_params(sub, qlum=ql, qchrom=qc, restart_interval=ri)
    cap = N * (W * H * 6 + 65536)
    out = np.empty(cap, np.uint8); offs = np.zeros(N, np.uint64); sizes = np.zeros(N, np.uint64)
    device = rng.random() < 0.5
    if device:
        d_in = enc.device_alloc(buf.nbytes); d_out = enc.device_alloc(cap); d_meta = enc.device_alloc(16 * N + 16)
        enc.h2d(d_in, buf)
        enc.encode_batch_device(d_in + lead, N, W, H, pitch, stride, p, d_out, cap, d_meta, d_meta + 8 * N, d_meta + 16 * N)
        enc.sync()
        enc.d2h(out, d_out); meta = np.zeros(2 * N + 1, np.uint64); enc.d2h(meta, d_meta)
        offs, sizes = meta[:N], meta[N: 2 * N]
        for d in (d_in, d_out, d_meta): enc.device_free(d)
    else:
        enc.encode_batch_ptr(buf.ctypes.data + lead, N, W, H, pitch, stride, p, out.ctypes.data, cap, offs, sizes)
    ok = True
    for f in range(N):
        jf = bytes(out[int(offs[f]): int(offs[f]) + int(sizes[f])])
        if jf != ol.encode_jfif(frames[f], sub, ql, qc, ri):
            ok = False; print("MISMATCH", c, f, sub, W, H, N, q, ri, pad, gap, lead, device); break
    bad += not ok
print("fuzz2 done:", cases, "cases,", bad, "bad")
