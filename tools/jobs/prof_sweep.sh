mkdir -p gpurun_out
for m in 1 3; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_sweep7 -c 1 -f -o gpurun_out/r2_sweep_v10_mode$m python tools/profile_sweep.py 2000 1100 1 $m > gpurun_out/r2_prof_mode$m.log 2>&1
tail -3 gpurun_out/r2_prof_mode$m.log
done
ls -la gpurun_out/*.ncu-rep
