#!/bin/bash
# 1 GPU: the tests behind (and including) the uncontracted-blend test, without -x
out=gpurun_out/${1:-recheck}; mkdir -p $out
( time timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_sort.py tests/test_host_c.py tests/test_refshaders.py tests/test_gpu_edges.py -m gpu -q ) > $out/pytest_a.log 2>&1
tail -6 $out/pytest_a.log
( time timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k "c3 or c2" ) > $out/pytest_b.log 2>&1
tail -4 $out/pytest_b.log
timeout 300 python bench.py --steps 60 --warmup 10 --no-radix > $out/bench.json 2> $out/bench.err; python -c "
import json; d=json.load(open('$out/bench.json')); print('fps',round(d['fps'],1),'e2e',round(d['e2e']['fps'],1),{k:round(v,3) for k,v in d['stage_ms'].items()}, d['parity'].get('rgba_bit_identical'))"
