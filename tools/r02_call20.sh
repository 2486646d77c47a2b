#!/bin/bash
# the whole GPU tier on the final tree, smoke(), per-kernel times on many small buffers
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3 | tee gpurun_out/gpu_tests_final2.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python tools/gpu_small_buffer_timing.py 2>&1 | tail -4
