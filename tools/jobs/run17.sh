mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest17.log 2>&1; tail -4 gpurun_out/r2_pytest17.log
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"^k_sweep" --csv --log-file gpurun_out/r2_sweep_v12.csv python tools/profile_sweep.py 2000 1100 1 > /dev/null 2>&1
grep -E "k_sweep" gpurun_out/r2_sweep_v12.csv | cut -d, -f5,13- | cut -c1-30,100- | head -6
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench17.json 2> gpurun_out/r2_bench17.err; tail -c 300 gpurun_out/r2_bench17.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench17.json'))
print('value',round(d['value'],1),'single ms',round(d['single_pair']['ms'],2),'e2e',round(d['e2e']['value'],1),'match',d['config']['e2e_outputs_match_device_run'])
"
