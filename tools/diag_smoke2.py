"""diagnostic: dirty the device memory with level 1-4 / inflate / crypto work in another process, then run smoke()'s deflate
workload in fresh processes and in a loop; report every mismatch (stream, first differing byte)"""
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
reps = int(sys.argv[1])
bufs = [datagen.silesia_mix(i, 40000 + 1000 * i).tobytes() for i in range(8)]
refs = [O.deflate(b, level=6) for b in bufs]
bad = 0
for rep in range(reps):
    outs, _ = z.deflate_batch(bufs, level=6)
    for i, (o, r) in enumerate(zip(outs, refs)):
        if o != r:
            bad += 1
            k = next((j for j in range(min(len(o), len(r))) if o[j] != r[j]), min(len(o), len(r)))
            print("  rep %%d stream %%d BAD at byte %%d of %%d (got %%d bytes)" %% (rep, i, k, len(r), len(o)), flush=True)
print("  %%d calls, %%d bad streams" %% (reps, bad), flush=True)
''' % (ROOT, ROOT)
DIRTY = r'''
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np
import sharpziplib_b200 as z
from sharpziplib_b200 import datagen
z.init(0)
bufs = [datagen.silesia_mix(i, 300000).tobytes() for i in range(64)]
for level in (1, 3, 0, 9):
    outs, _ = z.deflate_batch(bufs, level=level)
z.inflate_batch(outs, [len(b) for b in bufs])
import torch
x = torch.full((1 << 28,), -1, dtype=torch.int32, device="cuda"); del x
print("  dirtied", flush=True)
''' % (ROOT, ROOT)
for round_ in range(3):
    r = subprocess.run([sys.executable, "-c", DIRTY], capture_output=True, text=True, timeout=300)
    print(r.stdout[-200:], r.stderr[-300:], flush=True)
    r = subprocess.run([sys.executable, "-c", CHILD, "1" if round_ < 2 else "40"], capture_output=True, text=True, timeout=300)
    print(r.stdout[-1500:], r.stderr[-600:], flush=True)
