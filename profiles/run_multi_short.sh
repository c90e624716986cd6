#!/bin/bash
# Reduced multi-GPU refresh (8-GPU box; GPU-minutes are charged x8):  gpurun --gpus 8 -- 'bash profiles/run_multi_short.sh r02b'
R=${1:-r02b}
O=gpurun_out
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
timeout 200 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3 > $O/${R}_multi_test.log
timeout 300 bash -c "$(declare -f tr); tr 8 bench.py --gpus 8 --workload gigapixel --steps 5 --warmup 3" > $O/${R}_gigapixel_n8.json 2> $O/${R}_gigapixel_n8.err
timeout 300 bash -c "$(declare -f tr); tr 4 bench.py --gpus 4 --workload gigapixel --steps 5 --warmup 3 --no-e2e" > $O/${R}_gigapixel_n4.json 2> $O/${R}_gigapixel_n4.err
timeout 200 bash -c "$(declare -f tr); tr 8 bench.py --gpus 8 --workload strips16k --steps 10 --warmup 3" > $O/${R}_strips16k_n8.json 2> $O/${R}_strips16k_n8.err
timeout 300 bash -c "$(declare -f tr); tr 8 bench.py --gpus 8 --steps 10 --warmup 3" > $O/${R}_batch_n8.json 2> $O/${R}_batch_n8.err
for f in $O/${R}_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), d.get("stitch"), (d.get("parity_check") or {}).get("equal"))
except Exception as e:
    print("no result:", e)
PY
done
cat $O/${R}_multi_test.log
