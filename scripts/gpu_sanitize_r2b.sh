#!/bin/bash
out=gpurun_out/${1:-san3}; mkdir -p $out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 120 compute-sanitizer --tool memcheck --log-file $out/sanitizer_memcheck_group_G8.log python tests/group_inprocess_worker.py 8 60000 1280 720 1.2 1 > $out/group8.log 2>&1; tail -1 $out/group8.log; tail -2 $out/sanitizer_memcheck_group_G8.log
timeout 120 compute-sanitizer --tool racecheck --log-file $out/sanitizer_racecheck_group_G2.log python tests/group_inprocess_worker.py 2 30000 640 360 0.0 0 > $out/group2r.log 2>&1; tail -1 $out/group2r.log; tail -2 $out/sanitizer_racecheck_group_G2.log
