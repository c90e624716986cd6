# quick GPU loop: tensor-core parity tests, then the device-resident bench with either transform kernel
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tensor or tc" 2>&1 | tail -2
for t in ${MODES:-1 1 0}; do python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --tensor-dct $t | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($t, d['value'], d['roofline']['frac'], d['roofline']['step_breakdown_us'])"; done
