mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 tools/pcie_probe.py > gpurun_out/r2_pcie_probe_n8.json 2> gpurun_out/r2_pcie_probe_n8.err
echo "rc=$?"; tail -c 300 gpurun_out/r2_pcie_probe_n8.err
python -c "
import json
d=json.load(open('gpurun_out/r2_pcie_probe_n8.json'))
for r in d['results']: print(r['gpus_copying'], r['mode'], round(r['aggregate_gb_s'],1), round(r['per_gpu_gb_s'],1), round(r['ms_per_round'],1))
"
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" ; free -g | head -2
