mkdir -p gpurun_out
for b in 16 32; do
  timeout 400 python bench.py --steps 6 --warmup 3 --batch $b --no-cpu-baseline --no-stitch > gpurun_out/r2_bench13_b$b.json 2> gpurun_out/r2_bench13.err; tail -c 300 gpurun_out/r2_bench13.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench13_b$b.json'))
print('batch $b: value',round(d['value'],1),'sync',round(d['value_sync_api']['value'],1),'single ms',round(d['single_pair']['ms'],2),'e2e',round(d['e2e']['value'],1),'match',d['config']['e2e_outputs_match_device_run'], 'hbm frac', round(d['roofline']['whole_step_hbm_frac'],4))
"
done
