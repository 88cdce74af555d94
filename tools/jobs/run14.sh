mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest14.log 2>&1; tail -4 gpurun_out/r2_pytest14.log
python bench.py --workload stitch5 --steps 3 --warmup 1 > gpurun_out/r2_stitch5_b.json 2> gpurun_out/r2_stitch5_b.err; python -c "
import json
d=json.load(open('gpurun_out/r2_stitch5_b.json')); print('stitch5 s',d['value'],d['config']['per_iteration_ms'],d['vs_shipped_final_result'])"
python bench.py --workload four_input --steps 3 --warmup 1 > gpurun_out/r2_four_b.json 2>/dev/null; python -c "
import json
d=json.load(open('gpurun_out/r2_four_b.json')); print('four_input s',d['value'])"
timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench14.json'))
print('value',round(d['value'],1),'single ms',round(d['single_pair']['ms'],2),'e2e',round(d['e2e']['value'],1),'stitch',d['config']['stitch_iteration'])
"
