mkdir -p gpurun_out
for m in 2 1 0; do
  PF_SWEEP_MARGIN=$m timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench7.json'))
print('margin $m: value',round(d['value'],1),'single ms',round(d['single_pair']['ms'],2),'e2e',round(d['e2e']['value'],1))
"
done
python bench.py --workload four_input --crop95 --steps 3 --warmup 1 > gpurun_out/r2_four_input.json 2> gpurun_out/r2_four_input.err; python -c "
import json
d=json.load(open('gpurun_out/r2_four_input.json')); print('four_input crop95 s', d['value'], d['vs_shipped_final_result'])"
