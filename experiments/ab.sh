# same-box A/B of library builds: experiments/ab/*.so are swapped in turn into the package directory
for round in 1 2; do for v in ${VARIANTS:-A B}; do
  cp experiments/ab/$v.so jpeg-encoder-opencl_b200/libjpegb200.so
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --tensor-dct ${TC:-1} | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['value'], d['roofline']['frac'], d['roofline']['step_breakdown_us'])"
done; done
