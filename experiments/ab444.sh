# same-box A/B on the 4:4:4 workload and the CUDA-core 4:2:0 kernel
for v in ${VARIANTS:-A B}; do
  cp experiments/ab/$v.so jpeg-encoder-opencl_b200/libjpegb200.so
  for args in "--workload 4k444" "--tensor-dct 0"; do
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline $args | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', '$args', d['value'], d['roofline']['frac'], d['roofline']['step_breakdown_us'])"
done; done
