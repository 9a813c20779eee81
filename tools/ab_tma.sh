#!/bin/bash
# cp.async build vs TMA build of the dataflow kernel (IPMZ_DF_TMA), correctness then speed, one GPU call
export IPMZ_DATAFLOW_MIN_N=100
for n in 130 300 1000 3001; do echo -n "tma  "; IPMZ_DF_TMA=1 timeout 60 python tools/prof_factor.py $n 2 || echo "FAILED n=$n rc=$?"; done
unset IPMZ_DATAFLOW_MIN_N
for rep in 1 2; do for n in 8192 4096; do
  echo -n "cp.async "; timeout 100 python tools/prof_factor.py $n 5
  echo -n "tma      "; IPMZ_DF_TMA=1 timeout 100 python tools/prof_factor.py $n 5
done; done
