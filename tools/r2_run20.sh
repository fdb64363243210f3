#!/bin/bash
mkdir -p gpurun_out
for e in a b c d e; do
  timeout 300 python tools/debug_capture2.py $e 2>&1 | grep -v Warning | tail -4 | cut -c1-400
done
