"""Per-stage device timings of the host-to-host batch path (jb_encode_batch) with either transform kernel."""
import sys, time
sys.path.insert(0, '.')
import __graft_entry__ as g
jb = g.load()
import numpy as np, torch
enc = jb.Encoder()
N, W, H = 256, 1920, 1080
frames = jb.pinned_empty((N, H, W, 3))
d = torch.empty(H * W * 3, dtype=torch.uint8, device='cuda')
for i in range(N):
    enc.synth_device(100 + (i % 8), W, 0, H, W * 3, d.data_ptr())
    enc.sync()
    frames[i] = d.cpu().numpy().reshape(H, W, 3)
out = jb.pinned_empty((N * W * H // 2,))
offs = np.zeros(N, np.uint64); sizes = np.zeros(N, np.uint64)
for prof in (False, True):
    enc.set_profiling(prof)
    for fl in (0, jb.FLAG_FMA_DCT, 0, jb.FLAG_FMA_DCT):
        p = jb.make_params(jb.SUB_420, quality=75, flags=fl)
        enc.reset_counters()
        t0 = time.perf_counter()
        enc.encode_batch_ptr(frames.ctypes.data, N, W, H, W * 3, W * H * 3, p, out.ctypes.data, out.size, offs, sizes)
        t1 = time.perf_counter()
        tm = enc.timings()
        print('prof' if prof else 'plain', 'fma' if fl else 'tc ', round((t1 - t0) * 1e3, 2), 'ms',
              {k: round(v / 1e3, 2) for k, v in tm.items() if k.endswith('_us') and v}, int(tm['total_launches']))
