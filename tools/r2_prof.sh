#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"rf_(back|front)" -c 4 -o gpurun_out/prof_r2_rf3 python tools/profile_one.py ave 1 audio 160 > gpurun_out/ncu_r2_rf3.log 2>&1
tail -2 gpurun_out/ncu_r2_rf3.log
