#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_parity.py -m gpu -q -k "graphed or dropout or stack_matches" --tb=short -W ignore 2>&1 | tail -12 | cut -c1-250
timeout 600 python tools/host_overhead.py 2>&1 | tail -3 | cut -c1-1200
