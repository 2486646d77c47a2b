#!/bin/bash
set -u
O=gpurun_out/r02_c7
mkdir -p $O
python -m pytest tests -m gpu -q -x --timeout 900 --tb=short 2>&1 | tail -12 > $O/tests.log; tail -5 $O/tests.log
NI=256 REPS=5 python tools/prof_inflate.py > $O/inflate_text_parallel.json 2> $O/err1.txt; cat $O/inflate_text_parallel.json; tail -3 $O/err1.txt
NI=256 REPS=3 DATA=mix python tools/prof_inflate.py > $O/inflate_mix_parallel.json 2> $O/err3.txt; cat $O/inflate_mix_parallel.json
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 3000 $O/bench.json; tail -3 $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"; tail -c 800 $O/bench_ref.json
