mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_edge_cases.py tests/test_gpu_pipeline.py -x -q > gpurun_out/r2_pytest30.log 2>&1; tail -5 gpurun_out/r2_pytest30.log
for L in 8 2; do
  export PF_SWEEP_LANES=$L
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:"^k_sweep$" --csv --log-file gpurun_out/r2_sweep_lanes$L.csv python tools/profile_sweep.py 2000 1100 1 > /dev/null 2>&1
  echo "== lanes $L"; grep -E "k_sweep<" gpurun_out/r2_sweep_lanes$L.csv | cut -d, -f9,18- | tr '\n' ' '; echo
done
unset PF_SWEEP_LANES
for L in 8 2; do
  export PF_SWEEP_LANES_LATENCY=$L
  timeout 200 python bench.py --batch 1 --steps 4 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench30_l$L.json 2> gpurun_out/r2_bench30.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench30_l$L.json'))
print('latency lanes $L: single ms',round(d['single_pair']['ms'],2), 'value(b=1)', round(d['value'],1))
"
done
