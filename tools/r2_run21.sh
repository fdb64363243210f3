#!/bin/bash
timeout 300 python tools/debug_capture2.py f 2>&1 | grep -v Warning | tail -8 | cut -c1-400
