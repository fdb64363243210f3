#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowfused_gpu.py -q --tb=short --maxfail=12 -m gpu 2>&1 | tail -4 | cut -c1-200
for sh in "1 audio" "1 visual"; do
  timeout 300 python tools/time_kernels.py ave $sh 160 > gpurun_out/r2_time8_ave_$(echo $sh | tr ' ' '_').log 2>&1
  grep "^rf_" gpurun_out/r2_time8_ave_$(echo $sh | tr ' ' '_').log
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
cut -c1-300 gpurun_out/r2_bench_f.json
