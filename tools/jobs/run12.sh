mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q -k "async or strided or golden" > gpurun_out/r2_pytest12.log 2>&1; tail -2 gpurun_out/r2_pytest12.log
for dv in 1 2 3 4; do
  PF_SWEEP_CTA_DIVISOR=$dv timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; tail -c 200 gpurun_out/r2_bench12.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench12.json'))
print('divisor $dv: value',round(d['value'],1),'single ms',round(d['single_pair']['ms'],2),'e2e',round(d['e2e']['value'],1))
"
done
PF_SWEEP_CTA_DIVISOR=2 timeout 300 python bench.py --steps 4 --warmup 3 --batch 32 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench12.json'))
print('divisor 2 batch 32: value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))
"
