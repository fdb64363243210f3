#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "router or moments or bn or softmax or latent or gemm or attn" --tb=short -W ignore 2>&1 | tail -8 | cut -c1-250
for m in audio visual; do
  AVMOE_DUMP_TAGS=gpurun_out/r2_tags_3_$m.json timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/r2_warm_3_$m.csv python tools/profile_one.py ave 3 $m 160 > gpurun_out/r2_warm_3_$m.log 2>&1
  tail -1 gpurun_out/r2_warm_3_$m.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"xg_(post|ln_pre)_(bwd|fwd)" -c 8 -o gpurun_out/prof_r2_xg3 python tools/profile_one.py ave 3 audio 160 > gpurun_out/ncu_r2_xg3.log 2>&1
tail -1 gpurun_out/ncu_r2_xg3.log
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err
cut -c1-330 gpurun_out/r2_bench_h.json
AVMOE_GEMM_BN_BALANCE=0 timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_h0.json 2> gpurun_out/r2_bench_h0.err
cut -c1-330 gpurun_out/r2_bench_h0.json
