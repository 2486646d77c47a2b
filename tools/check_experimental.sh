#!/bin/bash
# Compiles sharpziplib_b200/csrc/experimental/*.cuh for sm_100a next to the kernels they would join (nothing is linked, the
# product build does not see them) and prints registers / shared memory / spills.
set -e
cd "$(dirname "$0")/../sharpziplib_b200/csrc"
cat > /tmp/b200z_experimental_check.cu <<EOT
#include "$(pwd)/b200z_deflate.cu"
namespace b200z {
#include "$(pwd)/experimental/k_tile_parse.cuh"
}
EOT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I"$(pwd)" -Xptxas -v -c /tmp/b200z_experimental_check.cu -o /tmp/b200z_experimental_check.o 2>&1 | grep -A2 "k_tile_parse" | head -8
cuobjdump --dump-resource-usage /tmp/b200z_experimental_check.o 2>/dev/null | grep -A1 k_tile_parse | head -3
