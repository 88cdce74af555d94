mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -k "sweep" -x -q > gpurun_out/r2_sweeptest.log 2>&1; tail -3 gpurun_out/r2_sweeptest.log
if grep -q "passed" gpurun_out/r2_sweeptest.log && ! grep -q "failed" gpurun_out/r2_sweeptest.log; then
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_sweep --csv --log-file gpurun_out/r2_sweep_v11.csv python tools/profile_sweep.py 2000 1100 2 > /dev/null 2>&1
  grep -E "gpu__time_duration|inst_executed" gpurun_out/r2_sweep_v11.csv | cut -d, -f5,13- | head -4
  run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench4_$tag.json 2> gpurun_out/r2_bench4_$tag.err; tail -c 200 gpurun_out/r2_bench4_$tag.err; python -c "
import json,sys
d=json.load(open('gpurun_out/r2_bench4_$tag.json'))
print('$tag: value',round(d['value'],1),'single ms',round(d['single_pair']['ms'],2),'e2e',round(d['e2e']['value'],1),'sweep frac',d['roofline']['frac'],'stitch',d['config']['stitch_iteration'])
"; }
  run default PF_DUMMY=1
  (time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest4.log 2>&1; tail -4 gpurun_out/r2_pytest4.log
fi
