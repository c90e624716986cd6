#!/bin/bash
# Same-box A/B of library builds over the default workload AND the other workloads the default run times
# (4k444, 8k, repl1080p, nv12_1080p).  Variants are built beforehand as profiles/ab/<name>.so (see ab.sh).
#   VARIANTS="A B" bash profiles/ab2.sh
for round in 1 2; do for v in ${VARIANTS:-A B}; do
  cp profiles/ab/$v.so jpeg-encoder-opencl_b200/libjpegb200.so
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-parity | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'batch1080p', d['value'], d['roofline']['frac'], d['roofline']['step_breakdown_us'])
for k,o in d['config'].get('other_workloads',{}).items():
    print('$v', k, o.get('value'), o.get('roofline_frac'), o.get('step_breakdown_us'), o.get('error'))"
done; done
