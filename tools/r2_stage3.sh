#!/bin/bash
for sh in "3 audio" "3 visual"; do
  for mc in 0 768; do
    AVMOE_RF_MAX_C=$mc timeout 300 python tools/time_kernels.py ave $sh 160 > gpurun_out/r2_s3_${mc}_$(echo $sh | tr ' ' '_').log 2>&1
    echo "== $sh rf_max_c=$mc"; grep -E "^rf_|all kernels|^xg_post|^xg_ln" gpurun_out/r2_s3_${mc}_$(echo $sh | tr ' ' '_').log | cut -c1-150
  done
done
