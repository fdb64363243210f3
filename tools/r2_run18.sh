#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "resample or tensor_cores or golden or stack_matches or hot_shapes" --tb=short -W ignore 2>&1 | tail -8 | cut -c1-250
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_m.json 2> gpurun_out/r2_bench_m.err
cut -c1-330 gpurun_out/r2_bench_m.json; tail -2 gpurun_out/r2_bench_m.err | cut -c1-200
