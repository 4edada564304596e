#!/bin/bash
# Builds the product library in-tree: winterfell_b200/libwinterfell_b200.so (sm_100a only).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unknown-pragmas --use_fast_math -ccbin /usr/bin/g++"
mkdir -p _build
pids=()
for f in ntt commit fri layout capi prover ${EXTRA_SRCS}; do
  if [ ! -f _build/$f.o ] || [ csrc/$f.cu -nt _build/$f.o ] || [ -n "$(find csrc include ../include -newer _build/$f.o \( -name '*.cuh' -o -name '*.hpp' -o -name '*.h' -o -name '*.inc' \) 2>/dev/null | head -1)" ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c csrc/$f.cu -o _build/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o libwinterfell_b200.so _build/*.o -lcudart -ccbin /usr/bin/g++
echo "built $(pwd)/libwinterfell_b200.so"
