#!/bin/bash
# Registers / shared memory / spills of the opt-in kernels of sharpziplib_b200/csrc/experimental/ (they are compiled into
# libb200z.so with b200z_deflate.cu and launched only with B200Z_TILE_PARSE=1|2), next to the kernels they would replace.
set -e
cd "$(dirname "$0")/../sharpziplib_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xptxas -v -c b200z_deflate.cu -o /tmp/b200z_experimental_check.o 2>&1 \
  | grep -A3 "Compiling entry function.*\(k_tile_parse\|k_parse_fix\|k_match\|k_parse_chunk\)" | grep -v "^--" | sed 's/_ZN5b200z[0-9]*//; s/EPK.*for/ for/'
