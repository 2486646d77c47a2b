#!/bin/bash
set -u
O=gpurun_out/r02_c4
mkdir -p $O
python -m pytest tests -m gpu -q -x --timeout 900 --tb=short 2>&1 | tail -40 > $O/tests.log; tail -15 $O/tests.log
NI=256 REPS=5 python tools/prof_inflate.py > $O/inflate_text_parallel.json 2> $O/err1.txt; cat $O/inflate_text_parallel.json; tail -3 $O/err1.txt
NI=256 REPS=3 DATA=mix python tools/prof_inflate.py > $O/inflate_mix_parallel.json 2> $O/err3.txt; cat $O/inflate_mix_parallel.json
timeout 600 python bench.py > $O/bench_flat.json 2> $O/bench_flat.err; echo "bench rc=$?"; tail -c 2500 $O/bench_flat.json; tail -5 $O/bench_flat.err
B200Z_MATCH=loop timeout 600 python bench.py --steps 5 > $O/bench_loop.json 2> $O/bench_loop.err; echo "bench loop rc=$?"; python - <<'PY'
import json
for f in ("bench_flat","bench_loop"):
    try:
        d=json.loads(open("gpurun_out/r02_c4/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], {k: round(v,3) for k,v in d["kernels_ms"].items()})
    except Exception as e: print(f, "ERR", e)
PY
