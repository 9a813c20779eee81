#!/bin/sh
# Builds ipm-zoo_b200/libipmz_b200.so for sm_100a (in-tree; the .so travels with gpurun).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${EXTRA_NVCC_FLAGS} -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++"
mkdir -p "$HERE/csrc/_obj"
SRCS="vector_kernels assemble full_system bunch_kaufman factor dataflow dataflow_tma trsv batch_fused solver linear_solvers"
HDRS="$HERE/csrc/ipmz_device.cuh $HERE/csrc/ldlt_device.cuh $HERE/csrc/ldlt_schedule.hpp $HERE/csrc/dataflow_kernel.cuh \
  $HERE/csrc/ipmz_kernels.h $HERE/csrc/vector_bodies.cuh $HERE/../include/ipmz.h"
pids=""
for f in $SRCS; do
  o="$HERE/csrc/_obj/$f.o"
  stale=0
  [ -f "$o" ] || stale=1
  for d in "$HERE/csrc/$f.cu" $HDRS; do
    if [ "$stale" = 0 ] && [ "$d" -nt "$o" ]; then stale=1; fi
  done
  if [ "$stale" = 1 ]; then
    rm -f "$o"  # a failed compile must not leave an old object for the link
    $NVCC $FLAGS ${VERBOSE:+-Xptxas -v} -c "$HERE/csrc/$f.cu" -o "$o" &
    pids="$pids $!"
  fi
done
for p in $pids; do
  wait "$p" || { echo "build.sh: a compile job failed" >&2; exit 1; }
done
OBJS=""
for f in $SRCS; do
  [ -f "$HERE/csrc/_obj/$f.o" ] || { echo "build.sh: missing object $f.o" >&2; exit 1; }
  OBJS="$OBJS $HERE/csrc/_obj/$f.o"
done
$NVCC -shared -ccbin /usr/bin/g++ -o "$HERE/libipmz_b200.so" $OBJS -lcudart_static -ldl -lrt -lpthread
echo "built $HERE/libipmz_b200.so"
# host-side C++ mirror of the reference interface (over the C ABI) + its demo
/usr/bin/g++ -std=c++17 -O2 -fPIC -shared -o "$HERE/libipmz_host.so" "$HERE/host/ipmz_numerical_optimization.cpp" \
  -L"$HERE" -lipmz_b200 -Wl,-rpath,'$ORIGIN'
/usr/bin/g++ -std=c++17 -O2 -o "$HERE/host/host_demo" "$HERE/host/host_demo.cpp" -L"$HERE" -lipmz_host -lipmz_b200 \
  -Wl,-rpath,'$ORIGIN/..'
/usr/bin/g++ -std=c++17 -O2 -pthread -o "$HERE/host/ipmz_cli" "$HERE/host/ipmz_cli.cpp" -L"$HERE" -lipmz_host -lipmz_b200 \
  -Wl,-rpath,'$ORIGIN/..'
echo "built $HERE/libipmz_host.so, host/host_demo and host/ipmz_cli"
