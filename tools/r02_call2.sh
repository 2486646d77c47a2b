#!/bin/bash
# gpurun call 2 of round 2: GPU test tier, inflate timings (parallel vs serial, text and mix), ncu --set full of the new kernels
set -u
O=gpurun_out/r02_c2
mkdir -p $O
python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4 > $O/tests.log; cat $O/tests.log
NI=256 REPS=5 python tools/prof_inflate.py > $O/inflate_text_parallel.json 2> $O/err1.txt; cat $O/inflate_text_parallel.json
NI=256 REPS=3 B200Z_INFLATE=serial python tools/prof_inflate.py > $O/inflate_text_serial.json 2> $O/err2.txt; cat $O/inflate_text_serial.json
NI=256 REPS=5 DATA=mix python tools/prof_inflate.py > $O/inflate_mix_parallel.json 2> $O/err3.txt; cat $O/inflate_mix_parallel.json
for k in k_dec1 k_resolve k_find k_dec2; do
  NI=128 REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k\$ --launch-skip 1 -c 1 \
     -o $O/ncu_$k python tools/prof_inflate.py > $O/ncu_$k.log 2>&1; echo "ncu $k rc=$?"
done
ls -la $O
