#!/bin/bash
# k_fast v2 (batched gathers, cursor replay): parity of the level 1-4 tests, timings, one ncu source-level capture at level 1
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_schedules.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "levels_0_to_4 or small_writes or plan_with_schedules or small_corpus or fast_and_stored or window_slides or preset_dictionary" 2>&1 | tail -3
timeout 400 python tools/gpu_fast_levels.py 1,3 --big 2>&1 | cut -c1-400 | tail -10
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_fast -s 1 -c 1 -f -o gpurun_out/k_fast_v2 python tools/prof_fast.py 2>&1 | tail -3
