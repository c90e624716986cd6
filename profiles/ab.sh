#!/bin/bash
# Same-box A/B of library builds, run on the GPU box (gpurun).  Variants are built HERE first, e.g.
#   make -C jpeg-encoder-opencl_b200 -j8 && cp jpeg-encoder-opencl_b200/libjpegb200.so profiles/ab/A.so
#   (change the source / add -D flags, rebuild) && cp ... profiles/ab/B.so
# profiles/ab/*.so are git-ignored but travel with the gpurun snapshot.  Each variant is swapped into the package
# directory in turn, twice, so that drift of the box shows up as a difference between the rounds.
#   VARIANTS="A B" WORKLOAD=batch1080p bash profiles/ab.sh
for round in 1 2; do for v in ${VARIANTS:-A B}; do
  cp profiles/ab/$v.so jpeg-encoder-opencl_b200/libjpegb200.so
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --workload ${WORKLOAD:-batch1080p} --tensor-dct ${TC:-1} | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['value'], d['roofline']['frac'], d['roofline']['step_breakdown_us'])"
done; done
