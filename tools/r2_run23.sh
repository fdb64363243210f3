#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowfused_gpu.py -m gpu -q -x -k "token" --tb=short -W ignore 2>&1 | tail -12 | cut -c1-300
timeout 900 python -m pytest tests -m gpu -q -x -k "latent or golden or adapter_gpu or graphed" --tb=short -W ignore 2>&1 | tail -6 | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_o.json 2> gpurun_out/r2_bench_o.err
cut -c1-330 gpurun_out/r2_bench_o.json | grep -o '"ms_per_step": [0-9.]*'; tail -2 gpurun_out/r2_bench_o.err | cut -c1-300
