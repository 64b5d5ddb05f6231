#!/bin/bash
# 1 GPU: compute-sanitizer memcheck over the code added late in round 2 (front/back overlap, capacity growth, in-process shard group)
out=gpurun_out/${1:-san2}; mkdir -p $out
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 170 compute-sanitizer --tool memcheck --log-file $out/sanitizer_memcheck_overlap.log python -m pytest tests/test_gpu_pipeline.py -m gpu -q -k "overlapped or capacity_grows or big_splats" > $out/pytest_overlap.log 2>&1; tail -2 $out/pytest_overlap.log; tail -3 $out/sanitizer_memcheck_overlap.log
timeout 170 compute-sanitizer --tool memcheck --log-file $out/sanitizer_memcheck_group.log python tests/group_inprocess_worker.py 2 30000 640 360 0.0 1 > $out/group.log 2>&1; tail -2 $out/group.log; tail -3 $out/sanitizer_memcheck_group.log
