#!/bin/bash
# levels 1-4: parity tests of the group-step k_fast on the B200, then its timings against the lane-0 statement
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_schedules.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "levels_0_to_4 or small_writes or plan_with_schedules or engine_state or small_corpus or fast_and_stored or window_slides or preset_dictionary or fuzz or roundtrip_like" 2>&1 | tail -5
timeout 600 python tools/gpu_fast_levels.py 1,2,3,4 --big 2>&1 | tail -20
