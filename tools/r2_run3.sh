#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowfused_gpu.py -q --tb=short --maxfail=12 2>&1 | tail -120 > gpurun_out/r2_run3_rf.log
tail -25 gpurun_out/r2_run3_rf.log
for sh in "1 audio" "1 visual" "3 audio" "3 visual"; do
  timeout 300 python tools/time_kernels.py ave $sh 160 > gpurun_out/r2_time_ave_$(echo $sh | tr ' ' '_').log 2>&1
done
head -12 gpurun_out/r2_time_ave_1_audio.log
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 --tb=line --deselect tests/test_rowfused_gpu.py -W ignore 2>&1 | tail -60 > gpurun_out/r2_run3_tests.log
tail -30 gpurun_out/r2_run3_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
cut -c1-300 gpurun_out/r2_bench_c.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rf_ -c 8 -o gpurun_out/prof_r2_rf python tools/profile_one.py ave 1 audio 160 > gpurun_out/ncu_r2_rf.log 2>&1
tail -3 gpurun_out/ncu_r2_rf.log
