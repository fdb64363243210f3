#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rf_tok|router_l1|router_bwd_dpool|bn_moments_sq|softmax_cols_kernel" -c 14 -o gpurun_out/prof_r2_new python tools/profile_one.py ave 1 audio 160 > gpurun_out/ncu_r2_new.log 2>&1
tail -2 gpurun_out/ncu_r2_new.log
