mkdir -p gpurun_out/r2b
./ubench/f32x2_latency > gpurun_out/r2b/f32x2_latency.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2b/pytest_gpu.log 2>&1
timeout 300 python ubench/trace_compositor.py 10 > gpurun_out/r2b/trace.log 2>&1; cp gpurun_out/trace.npy gpurun_out/r2b/trace_shipped.npy
GSR_COMP_V3=8 timeout 300 python ubench/trace_compositor.py 10 >> gpurun_out/r2b/trace.log 2>&1; cp gpurun_out/trace.npy gpurun_out/r2b/trace_v3_8.npy
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b/sanitizer_memcheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2b/sanitizer_memcheck_smoke.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b/sanitizer_racecheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2b/sanitizer_racecheck_smoke.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_sort.py -q -x > gpurun_out/r2b/sanitizer_memcheck_sort.log 2>&1; echo "rc=$?" >> gpurun_out/r2b/sanitizer_memcheck_sort.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_sort.py -q -x -k "not 2_25 and not property" > gpurun_out/r2b/sanitizer_racecheck_sort.log 2>&1; echo "rc=$?" >> gpurun_out/r2b/sanitizer_racecheck_sort.log
cat gpurun_out/r2b/f32x2_latency.txt; tail -15 gpurun_out/r2b/pytest_gpu.log; cat gpurun_out/r2b/trace.log; tail -4 gpurun_out/r2b/sanitizer_*.log
