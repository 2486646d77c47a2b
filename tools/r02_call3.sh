#!/bin/bash
set -u
O=gpurun_out/r02_c3
mkdir -p $O
python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4 > $O/tests.log; cat $O/tests.log
NI=256 REPS=5 python tools/prof_inflate.py > $O/inflate_text_parallel.json 2> $O/err1.txt; cat $O/inflate_text_parallel.json; tail -3 $O/err1.txt
NI=256 REPS=3 DATA=mix python tools/prof_inflate.py > $O/inflate_mix_parallel.json 2> $O/err3.txt; cat $O/inflate_mix_parallel.json
NI=256 REPS=2 DATA=mix B200Z_INFLATE=serial python tools/prof_inflate.py > $O/inflate_mix_serial.json 2> $O/err4.txt; cat $O/inflate_mix_serial.json
