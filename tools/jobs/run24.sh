mkdir -p gpurun_out
run() { # name batch
  timeout 400 python bench.py --steps 4 --warmup 3 --batch $2 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench24_$1.json 2> gpurun_out/r2_bench24.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench24_$1.json'))
print('$1 batch $2: value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'single ms',round(d['single_pair']['ms'],2))
"; }
run b24 24
run b32 32
export PF_LIB_PATH=$PWD/tools/jobs/libpf_min4.so
run min4_b16 16
run min4_b32 32
