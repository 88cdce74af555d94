mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r2f_pytest.log 2>&1; tail -4 gpurun_out/r2f_pytest.log
timeout 900 python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; tail -c 300 gpurun_out/r2f_bench_default.err
timeout 300 python bench.py --batch 1 --no-cpu-baseline --no-stitch > gpurun_out/r2f_bench_single_pair.json 2> /dev/null
timeout 300 python bench.py --workload stitch5 --steps 3 --warmup 1 > gpurun_out/r2f_bench_stitch5.json 2> gpurun_out/r2f_stitch5.err
timeout 300 python bench.py --workload four_input --crop95 --steps 3 --warmup 1 > gpurun_out/r2f_bench_four_input.json 2> /dev/null
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_4000x2000.csv python tools/one_pair.py 4000 2000 > gpurun_out/r2f_one_pair.log 2>&1
python tools/summarise_launches.py gpurun_out/r2f_launches_4000x2000.csv | head -8
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:^k_sweep$' -c 1 -f -o gpurun_out/r2f_sweep_v12 python tools/profile_sweep.py 2000 1100 1 > gpurun_out/r2f_prof_v12.log 2>&1; tail -1 gpurun_out/r2f_prof_v12.log
python -c "
import json
d=json.load(open('gpurun_out/r2f_bench_default.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'single',d['single_pair']['ms'],'cpu',d['cpu_baseline'],'clocks',d['clocks'])
for f in ('stitch5','four_input'):
    s=json.load(open('gpurun_out/r2f_bench_%s.json'%f)); print(f, s['value'], s.get('vs_shipped_final_result'))
"
