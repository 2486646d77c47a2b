#!/bin/bash
set -u
O=gpurun_out/r02_c6
mkdir -p $O
python -m pytest tests/test_gpu_inflate_parallel.py tests/test_gpu_restart.py -m gpu -q -x --timeout 900 --tb=short -k "not large_streams and not bench_c4" 2>&1 | tail -15 > $O/tests.log; tail -6 $O/tests.log
NI=256 REPS=5 python tools/prof_inflate.py > $O/inflate_text_parallel.json 2> $O/err1.txt; cat $O/inflate_text_parallel.json; tail -3 $O/err1.txt
NI=256 REPS=3 DATA=mix python tools/prof_inflate.py > $O/inflate_mix_parallel.json 2> $O/err3.txt; cat $O/inflate_mix_parallel.json
for k in k_resolve k_dec1; do
  NI=128 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k\$ --launch-skip 1 -c 1 \
     -o $O/ncu_$k python tools/prof_inflate.py > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
