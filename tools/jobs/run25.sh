mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_edge_cases.py -x -q > gpurun_out/r2_pytest25.log 2>&1; tail -2 gpurun_out/r2_pytest25.log
for rep in 1 2; do
for v in default noP; do
  if [ $v = default ]; then unset PF_LIB_PATH; else export PF_LIB_PATH=$PWD/tools/jobs/libpf_$v.so; fi
  timeout 200 python bench.py --batch 1 --steps 4 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench25_$v.json 2> gpurun_out/r2_bench25.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench25_$v.json'))
print('$v single ms',round(d['single_pair']['ms'],2), 'value(b=1)', round(d['value'],1))
"
done
done
