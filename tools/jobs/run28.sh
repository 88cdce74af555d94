mkdir -p gpurun_out
for b in 8 10; do
timeout 300 python bench.py --batch $b --steps 8 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench28_b$b.json 2> gpurun_out/r2_bench28.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench28_b$b.json'))
print('batch $b: value',round(d['value'],1),'e2e(async 2 slots)',round(d['e2e']['value'],1),'sync host',round(d['e2e']['sync_api_value'],1))
"
done
