#!/bin/bash
export IPMZ_DATAFLOW_MIN_N=100
for n in 130 200 300 1000 2048 3001; do timeout 120 python tools/prof_factor.py $n 2 || echo "FAILED n=$n rc=$?"; done
unset IPMZ_DATAFLOW_MIN_N
timeout 120 python tools/dbg_cfg2.py | grep -v trace
timeout 120 python tools/trsv_log.py 8192
timeout 120 python tools/prof_factor.py 8192 5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
