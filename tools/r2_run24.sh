#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_parity.py -m gpu -q -x -k "graphed" --tb=short -W ignore 2>&1 | tail -4 | cut -c1-300
AVMOE_DUMP_TAGS=gpurun_out/r2_tags24_1_audio.json timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2_warm24_1_audio.csv python tools/profile_one.py ave 1 audio 160 > gpurun_out/r2_warm24_1_audio.log 2>&1
AVMOE_DUMP_TAGS=gpurun_out/r2_tags24_3_audio.json timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2_warm24_3_audio.csv python tools/profile_one.py ave 3 audio 160 > gpurun_out/r2_warm24_3_audio.log 2>&1
