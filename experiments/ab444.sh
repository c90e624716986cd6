# same-box A/B of library builds on the 4K 4:4:4 q90 workload (8x8-MCU transform kernel)
for round in 1 2; do for v in ${VARIANTS:-A B}; do
  cp experiments/ab/$v.so jpeg-encoder-opencl_b200/libjpegb200.so
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --workload 4k444 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['value'], d['roofline']['frac'], d['roofline']['step_breakdown_us'])"
done; done
