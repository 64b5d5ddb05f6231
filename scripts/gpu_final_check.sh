#!/bin/bash
# 1 GPU, end of round: full GPU test tier, smoke, the default bench line, ncu launch list of the final kernels
out=gpurun_out/${1:-final}; mkdir -p $out
export CUDA_DEVICE_MAX_CONNECTIONS=32
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $out/pytest_gpu.log 2>&1
tail -4 $out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > $out/smoke.log 2>&1; tail -3 $out/smoke.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-radix > $out/ncu_bench.log 2>&1
python -c "
import json; d=json.load(open('$out/bench.json')); print('fps',round(d['fps'],1),'e2e',round(d['e2e']['fps'],1),{k:round(v,3) for k,v in d['stage_ms'].items()}, d['parity'].get('rgba_bit_identical'), d['roofline']['frac'], d['cpu_baseline']['value'], d['gpu_launches'])"
