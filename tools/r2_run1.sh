#!/bin/bash
# round-2 run 1: parity suite + bench + parity report + variant bench lines (1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -40 > gpurun_out/r2_run1_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
AVMOE_PARITY_TAG=_r2 timeout 600 python tools/parity_report.py > gpurun_out/r2_parity_report.log 2>&1
for v in avqa avvp s4 ms3; do
  timeout 600 python bench.py --variant $v --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_${v}_1gpu.json 2> gpurun_out/r2_bench_${v}_1gpu.err
done
tail -5 gpurun_out/r2_run1_tests.log
cut -c1-400 gpurun_out/r2_bench_a.json
