#!/bin/bash
for v in "" st4; do
  L=$PWD/ipm-zoo_b200/libipmz_b200.so; [ -n "$v" ] && L=$PWD/ipm-zoo_b200/ab/libipmz_$v.so
  for c in 110 220; do echo -n "${v:-base} dbg=2 count=$c: "; IPMZ_LIB=$L IPMZ_FUSED_DBG=2 timeout 300 python tools/prof_batched.py $c 2>&1 | tail -1; done
  echo -n "${v:-base} full 4096: "; IPMZ_LIB=$L timeout 300 python tools/prof_batched.py 4096 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -x -q -k "batch" 2>&1 | tail -2
