#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_launches_bench.log 2>&1
tail -c 300 gpurun_out/r2_launches_bench.log
wc -l gpurun_out/r2_launches.csv
