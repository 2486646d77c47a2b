#!/bin/bash
# k_match with spin rounds (kMatchSpin): parity of the level 5-9 tests, then the bench line (per-kernel times inside)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "small_corpus or strategies or window_slides or long_single or c3_shape or fuzz or input_after_flush or zlib_wrapper" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_spin.json 2> gpurun_out/bench_spin.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_spin.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"])
print(d.get("kernels_ms") or d.get("config",{}).get("kernels_ms") or [k for k in d.keys()])
PY
