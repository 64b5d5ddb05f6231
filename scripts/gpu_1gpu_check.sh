# 1-GPU check: full GPU test tier, compositor scheduling sweep (full frame and one rank's share of an 8-GPU group), schedule trace, bench line
out=gpurun_out/${1:-r2e}; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $out/pytest_gpu.log 2>&1
timeout 400 python ubench/compositor_sweep.py > $out/sweep.log 2>&1
timeout 400 python ubench/compositor_sweep.py 8 > $out/sweep_rows8.log 2>&1
timeout 300 python ubench/trace_compositor.py 10 > $out/trace.log 2>&1; cp gpurun_out/trace.npy $out/trace.npy
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err
tail -4 $out/pytest_gpu.log; cat $out/sweep.log $out/sweep_rows8.log | grep -v "^$"; cat $out/trace.log; python -c "
import json; d=json.load(open('$out/bench.json')); print('fps',d['fps'],'e2e',d['e2e']['fps'],d['e2e']['rgb32f_packed']['fps'],d['stage_ms'], d['parity'])"
