#!/bin/bash
# Multi-GPU measurements of round 2, run on an 8-GPU box:  gpurun --gpus 8 -- 'bash profiles/run_multi.sh r02'
# Strong scaling of config #5 (gigapixel) and of a 16384^2 image as RST strips at N = 1/2/4/8, the raw pinned-H2D fabric
# microbenchmark with 1/2/4/8 GPUs copying at once, the batch workload (config #4, weak scaling) at N = 8, the 2-GPU tests.
R=${1:-r02}
NMAX=${NMAX:-8}
O=gpurun_out
tr() { n=$1; shift; if [ "$n" = 1 ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; fi; }
nvidia-smi topo -m > $O/${R}_topo.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -5 > $O/${R}_multi_test.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29511 profiles/tools/h2d_fabric.py > $O/${R}_h2d_fabric.jsonl 2> $O/${R}_h2d_fabric.err
for n in 8 4 2 1; do
  [ $n -le $NMAX ] || continue
  timeout 300 bash -c "$(declare -f tr); tr $n bench.py --gpus $n --workload strips16k --steps 10 --warmup 3" > $O/${R}_strips16k_n$n.json 2> $O/${R}_strips16k_n$n.err
  timeout 600 bash -c "$(declare -f tr); tr $n bench.py --gpus $n --workload gigapixel --steps 5 --warmup 3" > $O/${R}_gigapixel_n$n.json 2> $O/${R}_gigapixel_n$n.err
done
timeout 300 bash -c "$(declare -f tr); tr $NMAX bench.py --gpus $NMAX --workload gigapixel --stitch nccl --steps 5 --warmup 3 --no-e2e" > $O/${R}_gigapixel_n${NMAX}_nccl.json 2> $O/${R}_gigapixel_n${NMAX}_nccl.err
timeout 300 bash -c "$(declare -f tr); tr $NMAX bench.py --gpus $NMAX --workload gigapixel --stitch peer-nccl --steps 5 --warmup 3 --no-e2e" > $O/${R}_gigapixel_n${NMAX}_peernccl.json 2> $O/${R}_gigapixel_n${NMAX}_peernccl.err
timeout 300 bash -c "$(declare -f tr); tr $NMAX bench.py --gpus $NMAX --steps 10 --warmup 3" > $O/${R}_batch_n$NMAX.json 2> $O/${R}_batch_n$NMAX.err
for f in $O/${R}_*.json; do echo "== $f"; tail -c 300 ${f%.json}.err | grep -v OMP_NUM | tail -3; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), d.get("stitch"), (d.get("parity_check") or {}).get("equal"))
except Exception as e:
    print("no result:", e)
PY
done
cat $O/${R}_h2d_fabric.jsonl; cat $O/${R}_multi_test.log
