#!/bin/bash
# Round profile capture, run on the GPU box (gpurun): each ncu pass follows a plain run of the same command that
# exited 0.  Outputs go to gpurun_out/ (scratch); profiles/summarize.py turns them into the committed summaries.
#   bash profiles/capture.sh r02
R=${1:-r02}
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-others --no-parity"
set -x
$B > gpurun_out/${R}_plain.json 2> gpurun_out/${R}_plain.err || exit 1
# launch list of the same command (cold-cache, serialised: compare shares, not absolutes)
# (the 512 k_synth launches that fill the input batch are filtered out by name)
ncu --metrics gpu__time_duration.sum --clock-control none -c 240 --csv --log-file gpurun_out/${R}_launches.csv \
    -k 'regex:k_transform|k_fixup|k_encode|k_scan|k_intervals|k_zero|k_pack|k_ff_count|k_int_out|k_finalize|k_stuff' $B > /dev/null 2>&1
if [ -z "$ONLY_NV12" ]; then  # (ONLY_NV12=1: launch list + the NV12 kernel only, the other captures are kept)
# ncu --set full, one complete step of the main kernels (default build: tcgen05 transform, cp.async staging)
ncu --set full --import-source on --clock-control none -k 'regex:k_transform_tc|k_fixup|k_encode|^k_pack$|k_ff_count|^k_stuff$' \
    --launch-skip 18 --launch-count 6 -f -o gpurun_out/${R}_prof_tc $B > /dev/null 2>&1
# (gpurun merges at most 64 MiB back: the two variant kernels below are captured only when asked for)
if [ -n "$CAPTURE_VARIANTS" ]; then
# the CUDA-core transform kernel (JB_FLAG_FMA_DCT)
ncu --set full --import-source on --clock-control none -k 'regex:^k_transform$' --launch-skip 3 --launch-count 1 -f \
    -o gpurun_out/${R}_prof_full $B --tensor-dct 0 > /dev/null 2>&1
# the TMA-staged variant of the 4:2:0 transform (JB_FLAG_TMA)
ncu --set full --import-source on --clock-control none -k 'regex:k_transform_tma' --launch-skip 3 --launch-count 1 -f \
    -o gpurun_out/${R}_prof_tma $B --tma > /dev/null 2>&1
fi
# the 8x8-MCU tcgen05 transform (4K 4:4:4 q90)
$B --workload 4k444 > gpurun_out/${R}_plain444.json 2>> gpurun_out/${R}_plain.err || exit 1
ncu --set full --import-source on --clock-control none -k 'regex:k_transform_tc3' --launch-skip 3 --launch-count 1 -f \
    -o gpurun_out/${R}_prof_tc3 $B --workload 4k444 > /dev/null 2>&1
# the K = 16 chroma contraction of the replicated 4:2:0 mode (the reference's own mode)
$B --workload repl1080p > gpurun_out/${R}_plainrepl.json 2>> gpurun_out/${R}_plain.err || exit 1
ncu --set full --import-source on --clock-control none -k 'regex:k_transform_tc3' --launch-skip 3 --launch-count 1 -f \
    -o gpurun_out/${R}_prof_tc3r $B --workload repl1080p > /dev/null 2>&1
fi
# the NV12-style input kernel: it runs among the default run's other workloads (first timed launch of it)
B2="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity"
$B2 > gpurun_out/${R}_plain_others.json 2>> gpurun_out/${R}_plain.err || exit 1
ncu --set full --import-source on --clock-control none -k 'regex:k_transform_tc_nv12' --launch-skip 3 --launch-count 1 -f \
    -o gpurun_out/${R}_prof_nv12 $B2 > /dev/null 2>&1
ls -la gpurun_out/${R}_*
