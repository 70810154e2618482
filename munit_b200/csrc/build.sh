#!/bin/bash
# Builds libmunit_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p build
for f in tapgemm simt heads api; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ ptx.cuh -nt build/$f.o ] || [ common.h -nt build/$f.o ] || [ ../../include/munit_b200.h -nt build/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o build/$f.o
  fi
done
$NVCC -shared -o ../libmunit_b200.so build/tapgemm.o build/simt.o build/heads.o build/api.o -cudart shared
echo built $(realpath ../libmunit_b200.so)
