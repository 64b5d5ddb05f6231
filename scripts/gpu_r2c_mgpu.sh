# usage: bash scripts/gpu_r2c_mgpu.sh <N>   (run under gpurun --gpus N)
N=${1:-2}
out=gpurun_out/r2c_${N}gpu; mkdir -p $out
nvidia-smi -L > $out/gpus.txt 2>&1
nvidia-smi topo -m > $out/topo.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_multi.py -q -x ) > $out/pytest_multi.log 2>&1
run() { # name, args...
  name=$1; shift
  NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" > $out/bench_$name.json 2> $out/bench_$name.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('$out/bench_$name.json')); print('fps',round(d['fps'],1),'e2e_fps',round(d['e2e']['fps'],1),'stages',{k:round(v,3) for k,v in d['stage_ms'].items()},'e2e_stages',{k:round(v,3) for k,v in d['e2e'].get('stage_ms',{}).items()},'host_ms',{k:round(v,3) for k,v in d['e2e'].get('host_enqueue_ms_per_step',{}).items()},'parity',d['parity'].get('rgba_bit_identical'))" 2>&1 | tail -1)"
}
run c3_default --steps 120 --warmup 10 --no-radix          # what the driver launches (group, rows-local read-back, overlap on), shorter
run c3_serial --steps 120 --warmup 10 --no-radix --overlap 0
run c4_default --steps 40 --warmup 10 --workload c4 --no-radix
run c4_serial --steps 40 --warmup 10 --workload c4 --no-radix --overlap 0
if [ "${2:-}" = "peer" ]; then
run c3_peer --steps 60 --warmup 10 --mgpu peer --no-radix
run c4_peer --steps 40 --warmup 10 --mgpu peer --workload c4 --no-radix
fi
if [ "${2:-}" = "root" ]; then run c3_group_root --steps 60 --warmup 10 --present root --no-radix; fi
tail -5 $out/pytest_multi.log
