#!/bin/bash
# ncu launch lists (gpu__time_duration per launch, serialised) of one batched cfg4 solve at 512 and 4096 problems:
# which kernels lose efficiency at the per-GPU share of the 8-GPU run
for c in 512 4096; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_batched$c.csv python tools/prof_batched.py $c > gpurun_out/ncu_b$c.log 2>&1
  tail -2 gpurun_out/ncu_b$c.log
done
