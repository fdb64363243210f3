#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv python tools/one_step.py ave 160 > gpurun_out/r2_launches_final.log 2>&1
tail -c 300 gpurun_out/r2_launches_final.log
wc -l gpurun_out/r2_launches_final.csv
