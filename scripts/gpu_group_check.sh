# 1-GPU: group (scatter) path validation with in-process ranks + full GPU tier + bench
out=gpurun_out/${1:-r2g}; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $out/pytest_gpu.log 2>&1
for G in 2 8; do CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 python ubench/group_inprocess_c3.py $G 6 > $out/group_inprocess_G$G.log 2>&1; done
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err
tail -6 $out/pytest_gpu.log; tail -4 $out/group_inprocess_G*.log; python -c "
import json; d=json.load(open('$out/bench.json')); print('fps',d['fps'],'e2e',d['e2e']['fps'],d['e2e']['rgb32f_packed']['fps'],d['stage_ms'],d['e2e'].get('host_enqueue_ms_per_step'))"
