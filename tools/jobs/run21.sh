mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest21.log 2>&1; tail -4 gpurun_out/r2_pytest21.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench21.json 2> gpurun_out/r2_bench21.err; tail -c 300 gpurun_out/r2_bench21.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench21.json'))
print('value',round(d['value'],1),'single ms',round(d['single_pair']['ms'],2),'e2e',round(d['e2e']['value'],1),'match',d['config']['e2e_outputs_match_device_run'])
"
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:^k_sweep$' -c 1 -f -o gpurun_out/r2_sweep_v12 python tools/profile_sweep.py 2000 1100 1 > gpurun_out/r2_prof_v12.log 2>&1; tail -2 gpurun_out/r2_prof_v12.log
