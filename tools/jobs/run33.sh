mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_edge_cases.py tests/test_gpu_pipeline.py -x -q > gpurun_out/r2_pytest33.log 2>&1; tail -2 gpurun_out/r2_pytest33.log
PF_LIB_PATH=$PWD/tools/jobs/libpf_sh.so timeout 200 python -m pytest tests/test_gpu_stages.py -k sweep -x -q 2>&1 | tail -1
for v in default sh none; do
  if [ $v = default ]; then unset PF_LIB_PATH; else export PF_LIB_PATH=$PWD/tools/jobs/libpf_$v.so; fi
  timeout 200 python bench.py --batch 1 --steps 5 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench33_$v.json 2> gpurun_out/r2_bench33.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench33_$v.json'))
print('$v single ms',round(d['single_pair']['ms'],2))
"
done
