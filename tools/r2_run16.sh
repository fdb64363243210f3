#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --dtype f32 > gpurun_out/r2_bench_f32_tf32.json 2> gpurun_out/r2_bench_f32_tf32.err
cut -c1-330 gpurun_out/r2_bench_f32_tf32.json; tail -2 gpurun_out/r2_bench_f32_tf32.err
for v in s4 ms3; do
  timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --variant $v --segments 5 > gpurun_out/r2_bench_${v}_bt80_1gpu.json 2> gpurun_out/r2_bench_${v}_bt80_1gpu.err
  cut -c1-330 gpurun_out/r2_bench_${v}_bt80_1gpu.json; tail -2 gpurun_out/r2_bench_${v}_bt80_1gpu.err
done
