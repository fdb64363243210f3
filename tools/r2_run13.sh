#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "router or moments or bn or softmax or latent or attn or expert or rows or xg or adapter" --tb=short -W ignore 2>&1 | tail -8 | cut -c1-250
for m in audio visual; do
  AVMOE_DUMP_TAGS=gpurun_out/r2_tags13_3_$m.json timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2_warm13_3_$m.csv python tools/profile_one.py ave 3 $m 160 > gpurun_out/r2_warm13_3_$m.log 2>&1
  tail -1 gpurun_out/r2_warm13_3_$m.log
done
AVMOE_DUMP_TAGS=gpurun_out/r2_tags13_1_audio.json timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2_warm13_1_audio.csv python tools/profile_one.py ave 1 audio 160 > gpurun_out/r2_warm13_1_audio.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_i.json 2> gpurun_out/r2_bench_i.err
cut -c1-330 gpurun_out/r2_bench_i.json
