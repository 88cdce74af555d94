mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stages.py -k "sweep" -x -q > gpurun_out/r2_memcheck_sweep.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_memcheck_sweep.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -k "golden or tiny" -x -q > gpurun_out/r2_memcheck_pipe.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_memcheck_pipe.log
