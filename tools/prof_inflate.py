"""Fixed inflate workload for timing and ncu captures: NI x 1 MiB text streams (config C2's shape) pre-deflated by the
oracle at level 6, one warm-up run, then REPS timed runs with the plan's per-kernel CUDA-event timings printed as JSON.
    NI=256 REPS=5 python tools/prof_inflate.py            (B200Z_INFLATE=serial for the serial kernel alone)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import oracle_lib as O  # noqa: E402
import sharpziplib_b200 as z  # noqa: E402
from sharpziplib_b200 import datagen  # noqa: E402

z.init(0)
ni, reps = int(os.environ.get("NI", "256")), int(os.environ.get("REPS", "5"))
kind = os.environ.get("DATA", "text")
if kind == "text":
    t_np = [datagen.text_buffer(i, 1 << 20, config=2) for i in range(ni)]
else:
    t_np = [datagen.silesia_mix(i, 1 << 20, config=3) for i in range(ni)]
comp = O.batch(0, [a.tobytes() for a in t_np], level=6, threads=16)
iplan = z.InflatePlan([len(c) for c in comp], [a.size for a in t_np])
h2 = np.zeros(iplan.in_bytes, dtype=np.uint8)
for o, c in zip(iplan.in_offsets, comp):
    h2[o:o + len(c)] = np.frombuffer(c, dtype=np.uint8)
i_in = torch.from_numpy(h2).cuda()
i_out = torch.empty(iplan.out_bytes, dtype=torch.uint8, device="cuda")
il = torch.zeros(ni, dtype=torch.int64, device="cuda")
ist = torch.zeros(ni, dtype=torch.int32, device="cuda")
iu = torch.zeros(ni, dtype=torch.int64, device="cuda")
iplan.set_timing(True)
acc = {}
tot = []
for r in range(reps + 1):
    i_out.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    iplan.run(i_in, i_out, il, ist, None, iu)
    e1.record()
    torch.cuda.synchronize()
    if r == 0:
        continue
    tot.append(e0.elapsed_time(e1))
    for k, v in iplan.timings().items():
        acc.setdefault(k, []).append(v)
out = i_out.cpu().numpy()
ok = int(ist.abs().sum()) == 0
for i in range(ni):
    o = iplan.out_offsets[i]
    if out[o:o + t_np[i].size].tobytes() != t_np[i].tobytes():
        ok = False
        print("MISMATCH stream", i, file=sys.stderr)
        break
U = sum(a.size for a in t_np)
C = sum(len(c) for c in comp)
ms = float(np.median(tot))
print(json.dumps({"streams": ni, "data": kind, "ok": ok, "U": U, "C": C, "ms": ms, "gbs_U": U / ms / 1e6, "gbs_CU": (U + C) / ms / 1e6,
                  "kernels_ms": {k: float(np.median(v)) for k, v in acc.items()}, "stats": iplan.stats(),
                  "mode": os.environ.get("B200Z_INFLATE", "parallel")}))
sys.exit(0 if ok else 1)
