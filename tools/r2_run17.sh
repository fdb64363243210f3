#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "moments or bn or expert or tf32_switch or golden" --tb=short -W ignore 2>&1 | tail -6 | cut -c1-250
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_l.json 2> gpurun_out/r2_bench_l.err
cut -c1-330 gpurun_out/r2_bench_l.json
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --dtype f32 > gpurun_out/r2_bench_f32_tf32.json 2> gpurun_out/r2_bench_f32_tf32.err
cut -c1-330 gpurun_out/r2_bench_f32_tf32.json
AVMOE_DUMP_TAGS=gpurun_out/r2_tags17_1_audio.json timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2_warm17_1_audio.csv python tools/profile_one.py ave 1 audio 160 > gpurun_out/r2_warm17_1_audio.log 2>&1
