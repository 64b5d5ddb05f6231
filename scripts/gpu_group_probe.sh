#!/bin/bash
# 1 GPU: in-process group correctness + the read-back probe on the ubench variant builds (ubench/_variants/libgsr_*.so)
mkdir -p gpurun_out/probe
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python -m pytest tests/test_gpu_group.py -x -q -m gpu 2>&1 | tail -3
for v in ubench/_variants/libgsr_*.so; do
  GSR_LIB_PATH=$PWD/$v timeout 200 python ubench/group_e2e_probe.py ${1:-2} 2>&1 | grep -v Warning | tee -a gpurun_out/probe/probe.log
done
