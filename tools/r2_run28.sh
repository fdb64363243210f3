#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short -W ignore 2>&1 | tail -4 | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
