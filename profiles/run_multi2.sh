O=gpurun_out
tr() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) "$@"; }
timeout 400 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -5 | tee $O/r2ab_multi_test.log
timeout 300 bash -c "$(declare -f tr); tr 2 bench.py --gpus 2 --workload strips16k --steps 10 --warmup 3" > $O/r2ab_strips16k_n2.json 2> $O/r2ab_strips16k_n2.err
timeout 600 bash -c "$(declare -f tr); tr 2 bench.py --gpus 2 --workload gigapixel --steps 5 --warmup 3" > $O/r2ab_gigapixel_n2.json 2> $O/r2ab_gigapixel_n2.err
timeout 300 bash -c "$(declare -f tr); tr 2 bench.py --gpus 2 --steps 10 --warmup 3" > $O/r2ab_batch_n2.json 2> $O/r2ab_batch_n2.err
for f in $O/r2ab_*.json; do echo "== $f"; tail -c 300 ${f%.json}.err | grep -v OMP_NUM | tail -3; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["n_gpus"], d["value"], d["ms_per_step"], (d.get("e2e") or {}).get("value"), d.get("stitch"), (d.get("parity_check") or {}).get("equal"))
except Exception as e:
    print("no result:", e)
PY
done
