mkdir -p gpurun_out/r2a
nvidia-smi -L > gpurun_out/r2a/gpus.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a/pytest_gpu.log 2>&1
for v in "X=0" GSR_COMP_V2=5 GSR_COMP_V2=6 GSR_COMP_V2=8 GSR_COMP_P4=1 GSR_COMP_V3=6 GSR_COMP_V3=8 "GSR_COMP_V3=6 GSR_COMP_CVT=1" "GSR_COMP_V3=8 GSR_COMP_CVT=1"; do timeout 300 env $v python ubench/compositor_variant.py >> gpurun_out/r2a/variants.log 2>&1; done
timeout 300 ./ubench/cub_sort > gpurun_out/r2a/cub_sort.json 2> gpurun_out/r2a/cub_sort.err
timeout 300 python ubench/sort_sweep.py 20 22 24 26 28 > gpurun_out/r2a/sort_sweep.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
tail -3 gpurun_out/r2a/pytest_gpu.log; cat gpurun_out/r2a/variants.log | grep -v "^$" | tail -20
