mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest9.log 2>&1; tail -6 gpurun_out/r2_pytest9.log
if ! grep -q "failed\|error" gpurun_out/r2_pytest9.log; then
for v in tma notma; do
  if [ $v = notma ]; then export PF_NO_TMA_TILES=1; else unset PF_NO_TMA_TILES; fi
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench9_$v.json 2> gpurun_out/r2_bench9_$v.err; tail -c 300 gpurun_out/r2_bench9_$v.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench9_$v.json'))
print('$v: value',round(d['value'],1),'single ms',round(d['single_pair']['ms'],2),'e2e',round(d['e2e']['value'],1))
"
done
fi
