#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "latent or attn or xg or rows or batch_sum" --tb=short -W ignore 2>&1 | tail -4 | cut -c1-250
timeout 600 python tools/gemm_kslope.py 2>&1 | tail -20
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_j.json 2> gpurun_out/r2_bench_j.err
cut -c1-330 gpurun_out/r2_bench_j.json
