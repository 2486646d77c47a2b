#!/bin/bash
# round 2, final one-GPU evidence: the whole GPU tier, both bench arms, config C4, the ncu launch list of the bench command and
# one full capture of the dominant kernel (k_match)
set -u
O=gpurun_out/r02_final
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4 | tee $O/gpu_tests.log
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?"
python - <<PY
import json
d=json.loads(open("$O/bench.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], {k: round(v,3) for k,v in d["kernels_ms"].items()})
r=json.loads(open("$O/bench_reference.json").read().strip().splitlines()[-1]); print("reference", r["value"], r["cpu_baseline"])
PY
timeout 900 python bench.py --config c4 > $O/c4.json 2> $O/c4.err; echo "c4 rc=$?"; cat $O/c4.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 > $O/bench_under_ncu.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_match --launch-skip 4 -c 1 -f -o $O/k_match_full python bench.py --steps 2 --warmup 1 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/k_match_full.ncu-rep --page raw --csv > $O/k_match_full_raw.csv 2>/dev/null
ls -la $O | tail -12
