#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowfused_gpu.py tests/test_optim.py -q --tb=short --maxfail=12 -m gpu 2>&1 | tail -80 > gpurun_out/r2_run4_rf.log
tail -8 gpurun_out/r2_run4_rf.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rf_(back|front)" -c 8 -o gpurun_out/prof_r2_rf2 python tools/profile_one.py ave 1 audio 160 > gpurun_out/ncu_r2_rf2.log 2>&1
tail -2 gpurun_out/ncu_r2_rf2.log
for sh in "1 audio" "3 visual"; do
  timeout 300 python tools/time_kernels.py ave $sh 160 > gpurun_out/r2_time_ave_$(echo $sh | tr ' ' '_').log 2>&1
done
head -14 gpurun_out/r2_time_ave_1_audio.log
timeout 900 python -m pytest tests/test_round2_parity.py tests/test_adapter_gpu.py -m gpu -q --maxfail=20 --tb=line -W ignore 2>&1 | tail -25 > gpurun_out/r2_run4_tests.log
tail -12 gpurun_out/r2_run4_tests.log | cut -c1-250
