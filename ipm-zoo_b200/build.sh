#!/bin/sh
# Builds ipm-zoo_b200/libipmz_b200.so for sm_100a (in-tree; the .so travels with gpurun).
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="${EXTRA_NVCC_FLAGS} -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++"
mkdir -p "$HERE/csrc/_obj"
for f in vector_kernels assemble full_system bunch_kaufman factor dataflow dataflow_tma trsv solver linear_solvers; do
  if [ ! -f "$HERE/csrc/_obj/$f.o" ] || [ "$HERE/csrc/$f.cu" -nt "$HERE/csrc/_obj/$f.o" ] || \
     [ "$HERE/csrc/ipmz_device.cuh" -nt "$HERE/csrc/_obj/$f.o" ] || [ "$HERE/csrc/ldlt_device.cuh" -nt "$HERE/csrc/_obj/$f.o" ] || \
     [ "$HERE/csrc/ldlt_schedule.hpp" -nt "$HERE/csrc/_obj/$f.o" ] || [ "$HERE/csrc/dataflow_kernel.cuh" -nt "$HERE/csrc/_obj/$f.o" ] || [ "$HERE/csrc/ipmz_kernels.h" -nt "$HERE/csrc/_obj/$f.o" ] || \
     [ "$HERE/../include/ipmz.h" -nt "$HERE/csrc/_obj/$f.o" ]; then
    $NVCC $FLAGS ${VERBOSE:+-Xptxas -v} -c "$HERE/csrc/$f.cu" -o "$HERE/csrc/_obj/$f.o" &
  fi
done
wait
$NVCC -shared -ccbin /usr/bin/g++ -o "$HERE/libipmz_b200.so" "$HERE"/csrc/_obj/*.o -lcudart_static -ldl -lrt -lpthread
echo "built $HERE/libipmz_b200.so"
# host-side C++ mirror of the reference interface (over the C ABI) + its demo
/usr/bin/g++ -std=c++17 -O2 -fPIC -shared -o "$HERE/libipmz_host.so" "$HERE/host/ipmz_numerical_optimization.cpp" \
  -L"$HERE" -lipmz_b200 -Wl,-rpath,'$ORIGIN'
/usr/bin/g++ -std=c++17 -O2 -o "$HERE/host/host_demo" "$HERE/host/host_demo.cpp" -L"$HERE" -lipmz_host -lipmz_b200 \
  -Wl,-rpath,'$ORIGIN/..'
/usr/bin/g++ -std=c++17 -O2 -pthread -o "$HERE/host/ipmz_cli" "$HERE/host/ipmz_cli.cpp" -L"$HERE" -lipmz_host -lipmz_b200 \
  -Wl,-rpath,'$ORIGIN/..'
echo "built $HERE/libipmz_host.so, host/host_demo and host/ipmz_cli"
