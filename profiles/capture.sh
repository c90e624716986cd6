#!/bin/bash
# Round profile capture, run on the GPU box (gpurun): each ncu pass follows a plain run of the same command that
# exited 0.  Outputs go to gpurun_out/ (scratch); profiles/summarize.py turns them into the committed summaries.
#   bash profiles/capture.sh r01
R=${1:-r01}
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
set -x
$B > gpurun_out/${R}_plain.json 2> gpurun_out/${R}_plain.err || exit 1
# launch list of the same command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $B > /dev/null 2>&1
# ncu --set full, one complete step of the main kernels (default build: tcgen05 transform)
ncu --set full --import-source on --clock-control none -k 'regex:k_transform_tc|k_fixup|k_encode|^k_pack$|k_ff_count|k_stuff' \
    --launch-skip 18 --launch-count 6 -f -o gpurun_out/${R}_prof_tc $B > /dev/null 2>&1
# the CUDA-core transform kernel (JB_FLAG_FMA_DCT)
ncu --set full --import-source on --clock-control none -k 'regex:k_transform<' --launch-skip 3 --launch-count 1 -f \
    -o gpurun_out/${R}_prof_full $B --tensor-dct 0 > /dev/null 2>&1
# the 8x8-MCU tcgen05 transform (4K 4:4:4 q90)
$B --workload 4k444 > gpurun_out/${R}_plain444.json 2>> gpurun_out/${R}_plain.err || exit 1
ncu --set full --import-source on --clock-control none -k 'regex:k_transform_tc3' --launch-skip 3 --launch-count 1 -f \
    -o gpurun_out/${R}_prof_tc3 $B --workload 4k444 > /dev/null 2>&1
ls -la gpurun_out/${R}_*
