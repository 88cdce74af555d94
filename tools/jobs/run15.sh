for lib in tools/jobs/libpf_min4.so tools/jobs/libpf_min5.so; do
  export PF_LIB_PATH=$PWD/$lib
  timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; tail -c 200 gpurun_out/r2_bench15.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench15.json'))
print('$lib value',round(d['value'],1),'single ms',round(d['single_pair']['ms'],2),'e2e',round(d['e2e']['value'],1))
"
done
