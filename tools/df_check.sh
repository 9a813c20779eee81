#!/bin/bash
# dataflow LDL^T: small-size correctness (incl. ragged tiles), then the n=8192 task log
export IPMZ_DATAFLOW_MIN_N=100
for n in 130 200 300 1000 2048 3001; do timeout 120 python tools/prof_factor.py $n 1 || echo "FAILED n=$n rc=$?"; done
unset IPMZ_DATAFLOW_MIN_N
timeout 300 python tools/df_tasklog.py 8192 gpurun_out/tasklog_8192.npy
timeout 300 python tools/df_tasklog.py 4096
timeout 120 python tools/prof_factor.py 8192 5
IPMZ_DATAFLOW_MIN_N=0 timeout 120 python tools/prof_factor.py 8192 5
