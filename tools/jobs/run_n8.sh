mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
(nvidia-smi dmon -s t -d 1 -c 40 > gpurun_out/r2_n8_dmon_pcie.txt 2>&1 &)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
echo "rc=$?"; tail -c 300 gpurun_out/r2_bench_8gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_8gpu.json'))
print('8gpu: value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'frac',round(d['e2e']['fraction_of_device_resident_value'],3),'sync api',round(d['e2e']['sync_api_value'],1),'e2e ms',d['e2e']['ms_per_step'])
"
grep -c . gpurun_out/r2_n8_dmon_pcie.txt
