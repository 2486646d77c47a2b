#!/bin/bash
# one GPU: the bench again (k_links run length), config C4 at its full 2 GiB, config C5 grid
set -u
O=gpurun_out/r02_c8
mkdir -p $O
timeout 600 python bench.py --steps 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('$O/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], {k: round(v,3) for k,v in d['kernels_ms'].items()})"
timeout 1500 python bench.py --config c4 > $O/c4.json 2> $O/c4.err; echo "c4 rc=$?"; cat $O/c4.json; tail -3 $O/c4.err
timeout 1500 python bench.py --config c5 > $O/c5.json 2> $O/c5.err; echo "c5 rc=$?"; python -c "
import json; d=json.loads(open('$O/c5.json').read().strip().splitlines()[-1]); print(d['all_parity']); [print(p['level'], p['size'], round(p['gbs'],2), round(p['ratio'],2), p['parity']) for p in d['points']]"; tail -3 $O/c5.err
