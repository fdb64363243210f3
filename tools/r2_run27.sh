#!/bin/bash
timeout 300 python tools/trace_small_ops.py 3 audio 2>&1 | tail -62 | cut -c1-200
