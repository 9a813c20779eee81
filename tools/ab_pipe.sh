OLD=$PWD/ipm-zoo_b200/ab/libipmz_b200_old.so
export IPMZ_DATAFLOW_MIN_N=100
for n in 130 1000 3001; do echo -n "new "; timeout 60 python tools/prof_factor.py $n 2 || echo FAILED; done
unset IPMZ_DATAFLOW_MIN_N
for rep in 1 2 3; do for n in 8192 4096; do
  echo -n "old: "; IPMZ_LIB=$OLD timeout 120 python tools/prof_factor.py $n 5
  echo -n "new: "; timeout 120 python tools/prof_factor.py $n 5
done; done
