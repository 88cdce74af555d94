mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -x -q > gpurun_out/r2_pytest16.log 2>&1; tail -3 gpurun_out/r2_pytest16.log
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_blur15|k_median5" --csv --log-file gpurun_out/r2_stencils_b.csv python tools/profile_stencils.py > /dev/null 2>&1
grep -E "k_blur15|k_median5" gpurun_out/r2_stencils_b.csv | grep -E "duration|inst_exec" | cut -d, -f5,13- | head -8
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench16.json 2> gpurun_out/r2_bench16.err; tail -c 200 gpurun_out/r2_bench16.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench16.json'))
print('value',round(d['value'],1),'single ms',round(d['single_pair']['ms'],2),'e2e',round(d['e2e']['value'],1))
"
