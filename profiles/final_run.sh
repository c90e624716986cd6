set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r02_tests.log
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_bench.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"
ONLY_NV12=${ONLY_NV12-1} bash profiles/capture.sh r02 > gpurun_out/r02_capture.log 2>&1; echo "capture rc=$?"
du -sh gpurun_out; ls -la gpurun_out/r02_*
