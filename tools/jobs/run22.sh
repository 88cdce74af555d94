mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_edge_cases.py tests/test_gpu_pipeline.py -x -q > gpurun_out/r2_pytest22.log 2>&1; tail -3 gpurun_out/r2_pytest22.log
for v in default noM noU noH none; do
  if [ $v = default ]; then unset PF_LIB_PATH; else export PF_LIB_PATH=$PWD/tools/jobs/libpf_$v.so; fi
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"^k_sweep$" --csv --log-file gpurun_out/r2_sweep_var_$v.csv python tools/profile_sweep.py 2000 1100 1 > /dev/null 2>&1
  echo "== $v"; grep -E "k_sweep<" gpurun_out/r2_sweep_var_$v.csv | cut -d, -f18- | tr '\n' ' '; echo
  timeout 200 python bench.py --batch 1 --steps 3 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench22_$v.json 2> gpurun_out/r2_bench22.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench22_$v.json'))
print('$v single ms',round(d['single_pair']['ms'],2))
"
done
