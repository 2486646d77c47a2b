#!/bin/bash
# k_fast v3 (persistent one-warp CTAs, head[] in a global pool): parity of the level 0-4 tests, timings
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_schedules.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "levels_0_to_4 or small_writes or plan_with_schedules or engine_state or small_corpus or fast_and_stored or window_slides or preset_dictionary or input_after_flush or handle_members or fuzz" 2>&1 | tail -3
timeout 400 python tools/gpu_fast_levels.py 1,2,3,4 --big 2>&1 | cut -c1-330 | tail -20
