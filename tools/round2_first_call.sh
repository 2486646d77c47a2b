#!/bin/bash
# One gpurun call that answers the questions round 1 left open, in the order that matters if it is cut short:
#   gpurun --timeout 2400 -- 'bash tools/round2_first_call.sh'      (about 30 GPU-minutes: tests 2, checks 2, three bench runs 9, five ncu passes 17)
# Everything lands in gpurun_out/r2_first/ (merged back by gpurun); copy what is to be judged into profiles/.
#   1. GPU test tier (is the product path still green; the opt-in kernels' subprocess test reports XPASS / xfail)
#   2. tools/tile_parse_check.py for the three opt-in search kernels: bit-exactness + per-kernel ms next to the default path
#   3. bench.py default, then with the fastest bit-exact variant (each line labels config.search_variant)
#   4. ncu: --set full of the opt-in kernel and of k_inflate / k_links / k_plan on tools/prof_small.py's workload, launch list of
#      the default bench
set -u
O=gpurun_out/r2_first
mkdir -p $O
python -m pytest tests -m gpu -q -x -rxX --timeout 600 > $O/tests.log 2>&1; echo "tests rc=$?" | tee -a $O/summary.txt
best=""; best_ms=1e9
for v in 3 4 2 1; do
  timeout 200 python tools/tile_parse_check.py $v 256 > $O/check_$v.json 2> $O/check_$v.err; rc=$?
  echo "tile_parse_check $v rc=$rc $(tail -n 1 $O/check_$v.json | cut -c1-600)" | tee -a $O/summary.txt
  if [ $rc -eq 0 ]; then
    ms=$(python -c "import json,sys; print(json.loads(open('$O/check_$v.json').read().strip().splitlines()[-1])['search_plus_parse_ms']['variant'])")
    if python -c "import sys; sys.exit(0 if $ms < $best_ms else 1)"; then best=$v; best_ms=$ms; fi
  fi
done
echo "fastest bit-exact variant: ${best:-none} ($best_ms ms search+parse on 256 x 256 KiB)" | tee -a $O/summary.txt
python bench.py --no-probe > $O/bench_default.json 2> $O/bench_default.err; echo "bench default rc=$?" | tee -a $O/summary.txt
if [ -n "$best" ]; then
  B200Z_LINK_RUN=131072 python bench.py --no-probe > $O/bench_link_run_128k.json 2> $O/bench_link_run_128k.err; echo "bench link run 128Ki rc=$?" | tee -a $O/summary.txt
  B200Z_TILE_PARSE=$best python bench.py > $O/bench_tile_parse$best.json 2> $O/bench_tile_parse$best.err; echo "bench variant $best rc=$?" | tee -a $O/summary.txt
  B200Z_TILE_PARSE=$best timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_parse -c 2 \
      -o $O/ncu_tile_parse$best python tools/prof_small.py > $O/ncu_tile_parse$best.log 2>&1; echo "ncu variant rc=$?" | tee -a $O/summary.txt
fi
for k in k_inflate k_links k_plan; do  # the next kernels in line once the search is faster (launch 0 is the warm-up run)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k --launch-skip 1 -c 1 \
      -o $O/ncu_$k python tools/prof_small.py > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?" | tee -a $O/summary.txt
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-probe > $O/launches_default.log 2>&1; echo "ncu launch list rc=$?" | tee -a $O/summary.txt
cat $O/summary.txt
