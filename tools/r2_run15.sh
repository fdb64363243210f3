#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "moments or bn or expert or adapter or tensor_cores" --tb=short -W ignore 2>&1 | tail -6 | cut -c1-250
AVMOE_DUMP_TAGS=gpurun_out/r2_tags15_1_audio.json timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2_warm15_1_audio.csv python tools/profile_one.py ave 1 audio 160 > gpurun_out/r2_warm15_1_audio.log 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err
cut -c1-330 gpurun_out/r2_bench_k.json
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --dtype f32 > gpurun_out/r2_bench_k_f32.json 2> gpurun_out/r2_bench_k_f32.err
cut -c1-330 gpurun_out/r2_bench_k_f32.json
