mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2_stitch_config3_launches.csv python tools/stitch_real_profile.py 1 > gpurun_out/r2_stitch_prof.log 2>&1
tail -2 gpurun_out/r2_stitch_prof.log
python tools/summarise_launches.py gpurun_out/r2_stitch_config3_launches.csv
