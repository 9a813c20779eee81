#!/bin/bash
mkdir -p gpurun_out
export IPMZ_FUSED_DBG=1 IPMZ_FUSED_CTAS_PER_SM=1
timeout 300 python tools/prof_batched.py 110 > gpurun_out/plain_mv.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ipm_batch -s 1 -c 1 -o gpurun_out/prof_mv -f python tools/prof_batched.py 110 > gpurun_out/ncu_mv.log 2>&1
tail -2 gpurun_out/plain_mv.log; tail -2 gpurun_out/ncu_mv.log
