#!/bin/bash
# usage: r2_multigpu.sh N "variant variant ..."   -- one torchrun bench line per variant at N GPUs (kept under gpurun_out/, copied to profiles/)
N=$1; shift
mkdir -p gpurun_out
if [ "$N" = "2" ] || [ "$N" = "4" ]; then
  timeout 600 python -m pytest tests/test_round2_parity.py -m gpu -q -k "dataparallel or non_current" --tb=short -W ignore 2>&1 | tail -15 > gpurun_out/r2_dp_tests.log
  tail -5 gpurun_out/r2_dp_tests.log
fi
for v in $@; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --variant $v --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_${v}_${N}gpu.json 2> gpurun_out/r2_bench_${v}_${N}gpu.err
  cut -c1-260 gpurun_out/r2_bench_${v}_${N}gpu.json
done
