mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_sweep<" -c 1 -f -o gpurun_out/r2_sweep_v11 python tools/profile_sweep.py 2000 1100 1 > gpurun_out/r2_prof_v11.log 2>&1; tail -1 gpurun_out/r2_prof_v11.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_blur15|k_median5" -c 2 -f -o gpurun_out/r2_stencils_tma python tools/profile_stencils.py > gpurun_out/r2_prof_st.log 2>&1; tail -1 gpurun_out/r2_prof_st.log
PF_NO_TMA_TILES=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_blur15|k_median5" --csv --log-file gpurun_out/r2_stencils_notma.csv python tools/profile_stencils.py > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_blur15|k_median5" --csv --log-file gpurun_out/r2_stencils_tma.csv python tools/profile_stencils.py > /dev/null 2>&1
grep -E "k_blur15|k_median5" gpurun_out/r2_stencils_notma.csv | cut -d, -f5,13- | head -8; grep -E "k_blur15|k_median5" gpurun_out/r2_stencils_tma.csv | cut -d, -f5,13- | head -8
timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -c 300 gpurun_out/r2_bench_default.err
timeout 300 python bench.py --batch 1 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench_single_pair.json 2> /dev/null
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_default.json')); r=json.load(open('gpurun_out/r2_bench_reference.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'single',d['single_pair'],'cpu',d['cpu_baseline'],'clocks',d['clocks'])
print('reference arm',r['value'],r['cpu_baseline'])
"
