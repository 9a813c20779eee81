#!/bin/bash
# fused batch kernel experiments: CTAs per SM, phase-only debug modes (dbg build)
C=${1:-4096}
for cps in 2 1; do
  echo "== product, CTAs/SM=$cps"; IPMZ_FUSED_CTAS_PER_SM=$cps timeout 300 python tools/prof_batched.py $C 2>&1 | tail -1
done
for cps in 2 1; do
for d in 1 2 3 4; do
  echo "== dbg mode $d, CTAs/SM=$cps"; IPMZ_FUSED_CTAS_PER_SM=$cps IPMZ_FUSED_DBG=$d IPMZ_LIB=$PWD/ipm-zoo_b200/ab/libipmz_dbg.so timeout 300 python tools/prof_batched.py $C 2>&1 | tail -1
done
done
