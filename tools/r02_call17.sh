#!/bin/bash
# parse with warm-up entry states: parity (long streams, C3 shape, 256 MiB, c4 small), then per-kernel times on large streams
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_inflate_parallel.py -m gpu -q -x --timeout 800 -k "small_corpus or window_slides or long_single or c3_shape or fuzz or input_after_flush or 256 or c4 or large or big" 2>&1 | tail -3
timeout 300 python tools/gpu_big_stream_timing.py 2>&1 | tail -4
