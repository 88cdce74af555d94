mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:^k_sweep$' -c 1 -f -o gpurun_out/r2_sweep_v11 python tools/profile_sweep.py 2000 1100 1 > gpurun_out/r2_prof_v11.log 2>&1; tail -3 gpurun_out/r2_prof_v11.log
