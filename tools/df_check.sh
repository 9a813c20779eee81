#!/bin/bash
export IPMZ_DATAFLOW_MIN_N=100
for n in 130 300 1000 3001; do timeout 120 python tools/prof_factor.py $n 2 || echo "FAILED n=$n rc=$?"; done
unset IPMZ_DATAFLOW_MIN_N
timeout 120 python tools/dbg_cfg2.py | grep -v trace
timeout 300 python tools/df_tasklog.py 8192 gpurun_out/tasklog_8192.npy | grep -v "DIAG start"
timeout 120 python tools/prof_factor.py 8192 5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
