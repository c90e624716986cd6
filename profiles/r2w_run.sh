set -x
cp profiles/ab/V.so jpeg-encoder-opencl_b200/libjpegb200.so
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?" 
tail -4 gpurun_out/r2w_tests.log
VARIANTS="U V" bash profiles/ab.sh 2>&1 | tee gpurun_out/r2w_ab.log
VARIANTS="U V" WORKLOAD=4k444 bash profiles/ab.sh 2>&1 | tee -a gpurun_out/r2w_ab.log
VARIANTS="U V" WORKLOAD=repl1080p bash profiles/ab.sh 2>&1 | tee -a gpurun_out/r2w_ab.log
cp profiles/ab/V.so jpeg-encoder-opencl_b200/libjpegb200.so
