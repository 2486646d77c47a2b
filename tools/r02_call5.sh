#!/bin/bash
set -u
O=gpurun_out/r02_c5
mkdir -p $O
NI=256 REPS=5 python tools/prof_inflate.py > $O/inflate_text_parallel.json 2> $O/err1.txt; cat $O/inflate_text_parallel.json; tail -3 $O/err1.txt
NI=256 REPS=3 DATA=mix python tools/prof_inflate.py > $O/inflate_mix_parallel.json 2> $O/err3.txt; cat $O/inflate_mix_parallel.json
for k in k_dec1 k_resolve k_find3 k_dec2; do
  NI=128 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k\$ --launch-skip 1 -c 1 \
     -o $O/ncu_$k python tools/prof_inflate.py > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
