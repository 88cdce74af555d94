mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_4000x2000.csv python tools/one_pair.py 4000 2000 > gpurun_out/r2_one_pair.log 2>&1
tail -2 gpurun_out/r2_one_pair.log
python tools/summarise_launches.py gpurun_out/r2_launches_4000x2000.csv
