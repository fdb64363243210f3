#!/bin/bash
# config 5 as specified: AVS S4 / MS3, 8 GPUs, 5 frames per clip -> BT = 80 per GPU
N=$1
mkdir -p gpurun_out
for v in s4 ms3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --variant $v --segments 5 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_${v}_bt80_${N}gpu.json 2> gpurun_out/r2_bench_${v}_bt80_${N}gpu.err
  grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/r2_bench_${v}_bt80_${N}gpu.json | head -2 | tr '\n' ' '; echo " $v"
done
