"""diagnostic: the smoke() workload and neighbours, per buffer, under a few settings (fresh process per setting)"""
import os
import subprocess
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import oracle_lib as O
import sharpziplib_b200 as z
from sharpziplib_b200 import datagen
z.init(0)
base = int(sys.argv[1])
bufs = [datagen.silesia_mix(i, base + 1000 * i).tobytes() for i in range(8)]
refs = [O.deflate(b, level=6) for b in bufs]
for rep in range(2):
    outs, _ = z.deflate_batch(bufs, level=6)
    res = []
    for b, o, r in zip(bufs, outs, refs):
        if o == r:
            res.append("ok")
        else:
            k = next((i for i in range(min(len(o), len(r))) if o[i] != r[i]), min(len(o), len(r)))
            res.append("BAD@%%d/%%d(%%d)" %% (k, len(r), len(o)))
    print("  call", rep, res, flush=True)
''' % (ROOT, ROOT)
for env, base in (({}, 40000), ({"B200Z_PARSE_WARM": "0"}, 40000), ({}, 30000), ({}, 70000), ({"B200Z_PARSE_WARM": "0"}, 70000)):
    print("env", env, "base", base, flush=True)
    r = subprocess.run([sys.executable, "-c", CHILD, str(base)], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
    print(r.stdout[-1500:], r.stderr[-600:], flush=True)
