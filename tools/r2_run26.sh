#!/bin/bash
mkdir -p gpurun_out
for v in avqa avvp s4 ms3; do
  timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager --variant $v --steps 5 > gpurun_out/r2_bench_${v}_1gpu.json 2> gpurun_out/r2_bench_${v}_1gpu.err
  echo "$v $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2_bench_${v}_1gpu.json | head -1) $(tail -1 gpurun_out/r2_bench_${v}_1gpu.err | cut -c1-200)"
done
