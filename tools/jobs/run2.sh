mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -k "sweep" -x -q > gpurun_out/r2_sweeptest.log 2>&1; tail -15 gpurun_out/r2_sweeptest.log
if grep -q "passed" gpurun_out/r2_sweeptest.log && ! grep -q "failed" gpurun_out/r2_sweeptest.log; then
  (time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest2.log 2>&1; tail -8 gpurun_out/r2_pytest2.log
  for m in 1 3; do
    timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_sweep7 --csv --log-file gpurun_out/r2_sweep_mode$m.csv python tools/profile_sweep.py 2000 1100 2 $m > /dev/null 2>&1
    grep -E "gpu__time_duration|inst_executed" gpurun_out/r2_sweep_mode$m.csv | cut -d, -f5,13- | head -8
  done
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -c 300 gpurun_out/r2_bench2.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench2.json'))
print('value',d['value'],'single',d['single_pair']['ms'],'e2e',d['e2e']['value'],'stitch',d['config']['stitch_iteration'])
"
fi
