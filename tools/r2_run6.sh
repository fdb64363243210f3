#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rowfused_gpu.py tests/test_optim.py -q --tb=short --maxfail=12 -m gpu 2>&1 | tail -60 > gpurun_out/r2_run6_rf.log
tail -6 gpurun_out/r2_run6_rf.log | cut -c1-200
for sh in "1 audio" "1 visual"; do
  timeout 300 python tools/time_kernels.py ave $sh 160 > gpurun_out/r2_time6_ave_$(echo $sh | tr ' ' '_').log 2>&1
  head -6 gpurun_out/r2_time6_ave_$(echo $sh | tr ' ' '_').log
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err
cut -c1-300 gpurun_out/r2_bench_e.json
timeout 1500 python -m pytest tests/test_adapter_gpu.py tests/test_config_sweep.py tests/test_round2_parity.py tests/test_stack_gpu.py -m gpu -q --maxfail=40 --tb=line -W ignore 2>&1 | tail -30 > gpurun_out/r2_run6_tests.log
tail -8 gpurun_out/r2_run6_tests.log | cut -c1-250
