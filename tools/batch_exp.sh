#!/bin/bash
for v in "" tma tmars; do
  L=$PWD/ipm-zoo_b200/libipmz_b200.so; [ -n "$v" ] && L=$PWD/ipm-zoo_b200/ab/libipmz_$v.so
  echo -n "${v:-base} full 4096: "; IPMZ_LIB=$L timeout 300 python tools/prof_batched.py 4096 2>&1 | tail -1
  echo -n "${v:-base} full 512: "; IPMZ_LIB=$L timeout 300 python tools/prof_batched.py 512 2>&1 | tail -1
done
