# 1-GPU profiling pass: group path per-kernel costs (in-process ranks), sparse-rule sweeps, ncu launch list + full-set captures of the bench
out=gpurun_out/${1:-r2f}; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > $out/pytest_gpu.log 2>&1
for G in 4 8; do
  CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/group_launches_G$G.csv python ubench/group_inprocess_c3.py $G 3 > $out/group_inprocess_G$G.log 2>&1
done
for rows in 1 2 4 8; do timeout 400 python ubench/compositor_sweep.py $rows > $out/sweep_rows$rows.log 2>&1; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-radix > $out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"composite_kernel|projection_kernel|onesweep_kernel|tile_ranges_kernel|tile_order_kernel|sort_hist_kernel" -s 40 -c 12 -o $out/prof_final -f python bench.py --steps 6 --warmup 4 --no-cpu-baseline --no-radix > $out/ncu_full.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench.json 2> $out/bench.err
tail -4 $out/pytest_gpu.log; cat $out/group_inprocess_G*.log | tail -14; grep -h BEST $out/sweep_rows*.log; python -c "
import json; d=json.load(open('$out/bench.json')); print('fps',d['fps'],'e2e',d['e2e']['fps'],d['e2e']['rgb32f_packed']['fps'],d['stage_ms'])"
