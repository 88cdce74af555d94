mkdir -p gpurun_out
for v in cw1 cw1p3 cw1p6 t0 tw1 tw1p3; do
  export PF_LIB_PATH=$PWD/tools/jobs/libpf_$v.so
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:"^k_sweep$" --csv --log-file gpurun_out/r2_sweep_var_$v.csv python tools/profile_sweep.py 2000 1100 1 > /dev/null 2>&1
  echo "== $v"; grep -E "k_sweep<" gpurun_out/r2_sweep_var_$v.csv | cut -d, -f18- | tr '\n' ' '; echo
  timeout 200 python bench.py --batch 1 --steps 3 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench19_$v.json 2> gpurun_out/r2_bench19.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench19_$v.json'))
print('$v single ms',round(d['single_pair']['ms'],2))
"
done
