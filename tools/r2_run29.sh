#!/bin/bash
timeout 900 python -m pytest tests/test_round2_parity.py -m gpu -q -k "batch_slice" --tb=short -W ignore 2>&1 | tail -12 | cut -c1-300
