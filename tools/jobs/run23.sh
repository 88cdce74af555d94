mkdir -p gpurun_out
for v in default p3 p5 p6 p4f3 p3f3; do
  if [ $v = default ]; then unset PF_LIB_PATH; else export PF_LIB_PATH=$PWD/tools/jobs/libpf_$v.so; fi
  timeout 200 python bench.py --batch 1 --steps 4 --warmup 3 --no-cpu-baseline --no-stitch > gpurun_out/r2_bench23_$v.json 2> gpurun_out/r2_bench23.err; python -c "
import json
d=json.load(open('gpurun_out/r2_bench23_$v.json'))
print('$v single ms',round(d['single_pair']['ms'],2), 'value(b=1)', round(d['value'],1))
"
done
