#!/bin/bash
# Builds libb200z.so in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall"
mkdir -p _build
for f in b200z_deflate b200z_inflate b200z_checksum b200z_api; do
  if [ ! -f _build/$f.o ] || [ $f.cu -nt _build/$f.o ] || [ b200z_core.cuh -nt _build/$f.o ] || [ b200z_internal.cuh -nt _build/$f.o ] || [ b200z_crc.cuh -nt _build/$f.o ] || [ experimental/k_tile_parse.cuh -nt _build/$f.o ] || [ ../../include/b200z.h -nt _build/$f.o ]; then
    $NVCC $FLAGS -c $f.cu -o _build/$f.o &
  fi
done
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../libb200z.so _build/b200z_deflate.o _build/b200z_inflate.o _build/b200z_checksum.o _build/b200z_api.o
echo built ../libb200z.so
