#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_round2_parity.py -m gpu -q -x -k "graphed" --tb=short -W ignore 2>&1 | tail -6 | cut -c1-300
timeout 600 python tools/host_overhead.py 2>&1 | tail -2 | cut -c1-1500
