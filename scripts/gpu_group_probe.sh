#!/bin/bash
# 1 GPU: in-process group correctness + the read-back probe on ubench variant builds (ubench/_variants/libgsr_*.so, git-ignored).
# Build them first, e.g. the per-CTA system fence that profiles/r02_group_e2e_probe.txt compares against:
#   python -c "from godotgaussiansplatting_b200 import build; build.build(force=True, extra=['-DGSR_SCATTER_FENCE_PER_CTA_SYS'], out='ubench/_variants/libgsr_ctasys.so')"
# (the scatter/wait/gather split of that log came from a probe build that has since been removed; the stage times remain)
mkdir -p gpurun_out/probe
export CUDA_DEVICE_MAX_CONNECTIONS=32
timeout 300 python -m pytest tests/test_gpu_group.py -x -q -m gpu 2>&1 | tail -3
for v in ubench/_variants/libgsr_*.so; do
  GSR_LIB_PATH=$PWD/$v timeout 200 python ubench/group_e2e_probe.py ${1:-2} 2>&1 | grep -v Warning | tee -a gpurun_out/probe/probe.log
done
