#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 --tb=line -W ignore 2>&1 | tail -25 > gpurun_out/r2_final_tests.log
tail -6 gpurun_out/r2_final_tests.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
cut -c1-400 gpurun_out/r2_bench_final.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
cut -c1-300 gpurun_out/r2_bench_reference_arm.json
