for t in 1 0; do python bench.py --workload strips16k --steps 5 --warmup 3 --no-cpu-baseline --tensor-dct $t | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($t, d['value'], d['ms_per_step'])"; done
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import __graft_entry__ as g
jb = g.load()
import torch
enc = jb.Encoder()
enc.set_profiling(True) if hasattr(enc, 'set_profiling') else None
W, rows = 16384, 8192
pitch = W * 3
d = torch.empty(rows * pitch, dtype=torch.uint8, device='cuda')
for y in range(0, rows, 1024):
    enc.synth_device(1, W, y, 1024, pitch, d.data_ptr() + y * pitch)
enc.sync()
out = torch.empty(rows * W // 2, dtype=torch.uint8, device='cuda')
for fl in (0, jb.FLAG_FMA_DCT):
    p = jb.make_params(jb.SUB_420, quality=75, restart_interval=1024, flags=jb.FLAG_CLAMP_SOF | fl)
    for it in range(3):
        enc.reset_counters() if hasattr(enc, 'reset_counters') else None
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = enc.encode_strip(d.data_ptr(), p, 0, False, W=W, rows=rows, pitch=pitch, device_io=True, out=out.data_ptr(), cap=out.numel())
        torch.cuda.synchronize(); t1 = time.perf_counter()
        tm = enc.timings()
        print(fl, it, round((t1 - t0) * 1e3, 3), n, {k: round(v, 1) for k, v in tm.items() if k.endswith('_us') and v})
PY
