#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowfused_gpu.py tests/test_optim.py -q --tb=short --maxfail=12 -m gpu 2>&1 | tail -30 > gpurun_out/r2_run5_rf.log
tail -4 gpurun_out/r2_run5_rf.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
cut -c1-300 gpurun_out/r2_bench_d.json
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --tb=line --deselect tests/test_rowfused_gpu.py -W ignore 2>&1 | tail -30 > gpurun_out/r2_run5_tests.log
tail -12 gpurun_out/r2_run5_tests.log | cut -c1-250
