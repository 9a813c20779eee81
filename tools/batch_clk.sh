#!/bin/bash
# batch parity tests, phase clocks of the fused batch kernel (instrumented build), product timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_indefinite.py -x -q -k "batch" > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_batch.log
tail -4 gpurun_out/pytest_batch.log
for c in ${COUNTS:-4096}; do
  echo "== clk build $c"; IPMZ_LIB=$PWD/ipm-zoo_b200/ab/libipmz_clk.so timeout 300 python tools/prof_batched.py $c 2>&1 | tail -3
  echo "== product $c"; timeout 300 python tools/prof_batched.py $c 2>&1 | tail -2
done
