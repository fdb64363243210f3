#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attn_aux_gpu.py tests/test_adapter_gpu.py tests/test_expert_rows_gpu.py -q --tb=line --maxfail=12 -m gpu -W ignore 2>&1 | tail -6 | cut -c1-200
timeout 300 python tools/time_kernels.py ave 1 audio 160 > gpurun_out/r2_time10_ave_1_audio.log 2>&1
grep -E "^softmax_cols |^bn_moments|all kernels" gpurun_out/r2_time10_ave_1_audio.log
timeout 300 python tools/time_kernels.py ave 3 audio 160 > gpurun_out/r2_time10_ave_3_audio.log 2>&1
grep -E "^softmax_cols |^bn_moments|all kernels" gpurun_out/r2_time10_ave_3_audio.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err
cut -c1-300 gpurun_out/r2_bench_g.json
