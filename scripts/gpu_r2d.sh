mkdir -p gpurun_out/r2d
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2d/pytest_gpu.log 2>&1
timeout 400 python ubench/compositor_sweep.py > gpurun_out/r2d/sweep_shipped.log 2>&1
GSR_COMP_V3=4 timeout 400 python ubench/compositor_sweep.py > gpurun_out/r2d/sweep_v3_4.log 2>&1
GSR_COMP_V3=4 GSR_COMP_PIPE=1 timeout 400 python ubench/compositor_sweep.py > gpurun_out/r2d/sweep_v3_4_pipe.log 2>&1
GSR_COMP_V3=6 timeout 400 python ubench/compositor_sweep.py > gpurun_out/r2d/sweep_v3_6.log 2>&1
timeout 400 python ubench/compositor_sweep.py 8 > gpurun_out/r2d/sweep_shipped_rows8.log 2>&1
GSR_COMP_V3=4 GSR_COMP_PIPE=1 timeout 400 python ubench/compositor_sweep.py 8 > gpurun_out/r2d/sweep_v3_4_pipe_rows8.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err
tail -4 gpurun_out/r2d/pytest_gpu.log; grep -h BEST gpurun_out/r2d/sweep_*.log; python -c "
import json; d=json.load(open('gpurun_out/r2d/bench.json')); print('fps',d['fps'],'e2e',d['e2e']['fps'],d['e2e']['rgb32f_packed']['fps'],d['stage_ms'])"
