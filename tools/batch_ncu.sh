#!/bin/bash
# one ncu --set full capture of the fused batch kernel (592 problems = two per resident CTA)
mkdir -p gpurun_out
C=${1:-592}
timeout 300 python tools/prof_batched.py $C > gpurun_out/plain_batch.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_ipm_batch -s 1 -c 1 -o gpurun_out/prof_batch -f python tools/prof_batched.py $C > gpurun_out/ncu_batch.log 2>&1
tail -3 gpurun_out/plain_batch.log; tail -3 gpurun_out/ncu_batch.log
