# ncu --set full capture (with source counters) of one launch of the tcgen05 transform kernel in the bench workload
ncu --set full --import-source on --clock-control none -k regex:k_transform_tc -s 3 -c 1 -f -o gpurun_out/${TAG:-tc}_prof \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG:-tc}_prof.log 2>&1
tail -3 gpurun_out/${TAG:-tc}_prof.log
