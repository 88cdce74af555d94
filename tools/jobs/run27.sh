mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_edge_cases.py tests/test_gpu_pipeline.py tests/test_cpp_mirror.py tests/test_stitch_prepare.py -x -q -m gpu > gpurun_out/r2_pytest27.log 2>&1; tail -3 gpurun_out/r2_pytest27.log
for rep in 1 2; do
for v in on off; do
  if [ $v = on ]; then unset PF_NO_FRONT_OVERLAP; else export PF_NO_FRONT_OVERLAP=1; fi
  timeout 200 python bench.py --batch 1 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench27_$v.json 2> gpurun_out/r2_bench27.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench27_$v.json'))
print('overlap $v: single ms',round(d['single_pair']['ms'],2), 'value(b=1)', round(d['value'],1), 'stitch', d['config']['stitch_iteration'])
"
done
done
