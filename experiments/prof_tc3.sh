# ncu --set full capture (with source counters) of one launch of the 8x8-MCU tcgen05 transform kernel (4K 4:4:4 q90)
ncu --set full --import-source on --clock-control none -k regex:k_transform_tc3 -s 3 -c 1 -f -o gpurun_out/${TAG:-tc3}_prof \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload 4k444 > gpurun_out/${TAG:-tc3}_prof.log 2>&1
tail -2 gpurun_out/${TAG:-tc3}_prof.log | cut -c1-300
