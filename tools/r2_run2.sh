#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowfused_gpu.py -q --tb=short -x 2>&1 | tail -60 > gpurun_out/r2_run2_rf.log
tail -3 gpurun_out/r2_run2_rf.log
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 --tb=line --deselect tests/test_rowfused_gpu.py 2>&1 | tail -80 > gpurun_out/r2_run2_tests.log
tail -45 gpurun_out/r2_run2_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
cut -c1-300 gpurun_out/r2_bench_b.json
