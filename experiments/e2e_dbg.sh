# e2e through jb_encode_batch with either transform kernel, and the raw pinned H2D bandwidth of the box
for t in 1 0; do python bench.py --steps 5 --warmup 3 --no-cpu-baseline --tensor-dct $t | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($t, d['value'], d['e2e']['value'], d['e2e']['ms_per_step'])"; done
python - <<'PY'
import torch, time
n = 3185049600
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); t1 = time.perf_counter()
print('raw pinned H2D GB/s', 3 * n / (t1 - t0) / 1e9)
PY
