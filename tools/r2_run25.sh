#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_parity.py tests/test_rowfused_gpu.py -m gpu -q -k "graphed or token" --tb=short -W ignore 2>&1 | tail -8 | cut -c1-400
for cfg in "0 1024" "1 1024" "1 16" "0 1024" "1 1024"; do
  set -- $cfg
  AVMOE_RF_TOKENS=$1 AVMOE_RF_TOK_MIN_NK=$2 timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1 | sed "s/^/tokens=$1 min_nk=$2 /"
done
