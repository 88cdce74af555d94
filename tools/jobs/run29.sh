mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -k "blur or diffusion" -x -q > gpurun_out/r2_pytest29.log 2>&1; tail -2 gpurun_out/r2_pytest29.log
for v in default adjcap3 noadj; do
  if [ $v = default ]; then unset PF_LIB_PATH; else export PF_LIB_PATH=$PWD/tools/jobs/libpf_$v.so; fi
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_blur15" --csv --log-file gpurun_out/r2_blur_var_$v.csv python tools/profile_stencils.py > /dev/null 2>&1
  echo "== $v"; grep -E "k_blur15" gpurun_out/r2_blur_var_$v.csv | cut -d, -f13- | tr '\n' ' '; echo
done
for v in default adjcap3 noadj; do
  if [ $v = default ]; then unset PF_LIB_PATH; else export PF_LIB_PATH=$PWD/tools/jobs/libpf_$v.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench29_$v.json 2> gpurun_out/r2_bench29.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench29_$v.json'))
print('$v value',round(d['value'],1),'single ms',round(d['single_pair']['ms'],2))
"
done
