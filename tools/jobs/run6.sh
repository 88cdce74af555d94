mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_reference_data.py -x -q) > gpurun_out/r2_refdata.log 2>&1; tail -15 gpurun_out/r2_refdata.log
python bench.py --workload stitch5 --steps 3 --warmup 1 --save-result gpurun_out/r2_stitch5_final.png > gpurun_out/r2_stitch5.json 2> gpurun_out/r2_stitch5.err; cat gpurun_out/r2_stitch5.json; tail -c 300 gpurun_out/r2_stitch5.err
python bench.py --workload four_input --crop95 --steps 3 --warmup 1 > gpurun_out/r2_four_input.json 2> gpurun_out/r2_four_input.err; cat gpurun_out/r2_four_input.json; tail -c 300 gpurun_out/r2_four_input.err
python bench.py --workload four_input --steps 3 --warmup 1 > gpurun_out/r2_four_input_nocrop.json 2> gpurun_out/r2_four_input_nocrop.err; cat gpurun_out/r2_four_input_nocrop.json
