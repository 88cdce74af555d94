mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 --redirects 3 --tee 3 --log-dir gpurun_out/trlogs bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "rc=$?"
tail -c 1500 gpurun_out/r2_bench_2gpu.err
find gpurun_out/trlogs -type f | head; for f in $(find gpurun_out/trlogs -name "stderr.log"); do echo "== $f"; tail -c 1500 $f; done
head -c 600 gpurun_out/r2_bench_2gpu.json
