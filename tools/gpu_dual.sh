#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large_parity.py -x -q -k "dual or 3- or -3 or DUAL or golden or trace" 2>&1 | tail -12
