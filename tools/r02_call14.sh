#!/bin/bash
# config C5 on one GPU with the group-step k_fast; source-level ncu capture of k_match on 128 x 256 KiB
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --config c5 > gpurun_out/c5_1gpu.json 2> gpurun_out/c5_1gpu.err; echo "c5 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c5_1gpu.json").read().strip().splitlines()[-1]); print("c5", d["n_gpus"], d["all_parity"], [(p["level"],p["size"],round(p["gbs"],2)) for p in d["points"]])
except Exception as e: print("c5 ERR", e)
PY
tail -2 gpurun_out/c5_1gpu.err
ND=128 NI=4 timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_match -s 1 -c 1 -f -o gpurun_out/k_match_r02 python tools/prof_small.py 2>&1 | tail -2
