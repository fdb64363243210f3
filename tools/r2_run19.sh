#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "resample or expert or rows or xg or adapter_gpu or stage" --tb=short -W ignore 2>&1 | tail -6 | cut -c1-250
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_n$i.json 2> gpurun_out/r2_bench_n$i.err
cut -c1-330 gpurun_out/r2_bench_n$i.json | grep -o '"ms_per_step": [0-9.]*'
done
