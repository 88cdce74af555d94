mkdir -p gpurun_out
export PF_SWEEP_LANES_LATENCY=8
for v in w8 w2 w8p2; do
  export PF_LIB_PATH=$PWD/tools/jobs/libpf_$v.so
  timeout 200 python bench.py --batch 1 --steps 4 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench31_$v.json 2> gpurun_out/r2_bench31.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench31_$v.json'))
print('lanes 8 $v: single ms',round(d['single_pair']['ms'],2))
"
done
