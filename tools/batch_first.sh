#!/bin/bash
# first GPU pass of the fused batch kernel: batch parity tests, then old-vs-fused timing at 512 / 4096 problems
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_indefinite.py -x -q -k "batch" > gpurun_out/pytest_batch.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_batch.log
tail -15 gpurun_out/pytest_batch.log
for c in 512 4096; do
  echo "== fused $c"; timeout 300 python tools/prof_batched.py $c 2>&1 | tail -2
  echo "== old $c"; IPMZ_BATCH_FUSED=0 timeout 300 python tools/prof_batched.py $c 2>&1 | tail -2
done
echo "== fused aug 4096"; timeout 300 python tools/prof_batched.py 4096 aug 2>&1 | tail -2
