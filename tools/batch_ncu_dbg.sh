#!/bin/bash
# ncu capture of one phase-only mode: batch_ncu_dbg.sh <dbg> <count>
mkdir -p gpurun_out
export IPMZ_FUSED_DBG=$1
timeout 300 python tools/prof_batched.py $2 > gpurun_out/plain_dbg.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ipm_batch -s 1 -c 1 -o gpurun_out/prof_dbg$1 -f python tools/prof_batched.py $2 > gpurun_out/ncu_dbg.log 2>&1
tail -1 gpurun_out/plain_dbg.log; tail -1 gpurun_out/ncu_dbg.log
