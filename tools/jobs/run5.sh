mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_stages.py -k "sweep" -x -q > gpurun_out/r2_sweeptest.log 2>&1; tail -2 gpurun_out/r2_sweeptest.log
for lib in default tools/jobs/libpf_s2_n4.so tools/jobs/libpf_s2_n6.so tools/jobs/libpf_s1_n8.so; do
  if [ "$lib" = default ]; then unset PF_LIB_PATH; else export PF_LIB_PATH=$PWD/$lib; fi
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_sweep --csv --log-file gpurun_out/r2_sweep_var.csv python tools/profile_sweep.py 2000 1100 2 > /dev/null 2>&1
  echo "== $lib"; grep -E "k_sweep<" gpurun_out/r2_sweep_var.csv | grep -E "gpu__time_duration|inst_executed" | cut -d, -f13- | head -4 | tr '\n' ' '; echo
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench5.json'))
print('   value',round(d['value'],1),'single ms',round(d['single_pair']['ms'],2),'e2e',round(d['e2e']['value'],1))
"
done
