#!/bin/bash
# 1 GPU: front/back overlap -- correctness (pipeline + in-process group tests) and the A/B measurement
out=gpurun_out/${1:-ab}; mkdir -p $out
export CUDA_DEVICE_MAX_CONNECTIONS=32
( time timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_group.py -m gpu -q -x ) > $out/pytest.log 2>&1
tail -5 $out/pytest.log
timeout 300 python ubench/pipeline_ab.py 200 2>&1 | grep -v Warning | tee $out/pipeline_ab.log
timeout 200 python ubench/group_e2e_probe.py 2 2>&1 | grep -v Warning | tee $out/group_probe.log
