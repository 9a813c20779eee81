#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_large_parity.py -x -q -k "late" 2>&1 | tail -3
IPMZ_CREATE_TIMING=1 timeout 600 python tools/create_timing.py 2>&1 | tail -16
