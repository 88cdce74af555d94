N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r2f_bench_${N}gpu.json 2> gpurun_out/r2f_bench_${N}gpu.err
echo "rc=$?"; tail -c 300 gpurun_out/r2f_bench_${N}gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r2f_bench_${N}gpu.json'))
print('${N}gpu: value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['e2e']['fraction_of_device_resident_value'],3),'e2e ms',round(d['e2e']['ms_per_step'],1))
"
