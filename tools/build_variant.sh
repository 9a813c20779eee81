#!/bin/sh
# tools/build_variant.sh NAME "EXTRA NVCC FLAGS" [file ...]: a second build of the library for A/B runs inside one GPU call
# (IPMZ_LIB=ipm-zoo_b200/ab/libipmz_NAME.so).  Only the listed .cu files (default: batch_fused) are recompiled with the
# extra flags; the other objects are the product build's.
set -e
HERE=$(cd "$(dirname "$0")/.." && pwd)/ipm-zoo_b200
NAME=$1; FLAGS_X=$2; shift 2 || true
FILES=${*:-batch_fused}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++"
mkdir -p "$HERE/ab/_obj_$NAME"
OBJS=""
for o in "$HERE"/csrc/_obj/*.o; do
  b=$(basename "$o" .o); skip=0
  for f in $FILES; do [ "$b" = "$f" ] && skip=1; done
  [ $skip = 1 ] || OBJS="$OBJS $o"
done
for f in $FILES; do
  $NVCC $FLAGS $FLAGS_X -c "$HERE/csrc/$f.cu" -o "$HERE/ab/_obj_$NAME/$f.o"
  OBJS="$OBJS $HERE/ab/_obj_$NAME/$f.o"
done
$NVCC -shared -ccbin /usr/bin/g++ -o "$HERE/ab/libipmz_$NAME.so" $OBJS -lcudart_static -ldl -lrt -lpthread
echo "built $HERE/ab/libipmz_$NAME.so"
