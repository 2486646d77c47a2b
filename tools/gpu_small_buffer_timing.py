"""per-kernel times of level-6 deflate plans over many small buffers (the small end of C5)"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import sharpziplib_b200 as z  # noqa: E402
from sharpziplib_b200 import datagen  # noqa: E402

z.init(0)
rows = []
for size, nbuf in ((4096, 65536), (16384, 16384), (65536, 4096)):
    base = [datagen.silesia_mix(i, size, config=5) for i in range(64)]
    plan = z.DeflatePlan([size] * nbuf, level=6)
    h = np.zeros(plan.in_bytes, np.uint8)
    for i, o in enumerate(plan.in_offsets):
        h[o:o + size] = base[i % 64]
    din = torch.from_numpy(h).cuda()
    dout = torch.empty(plan.out_bytes, dtype=torch.uint8, device="cuda")
    dl = torch.zeros(nbuf, dtype=torch.int64, device="cuda")
    ds = torch.zeros(nbuf, dtype=torch.int32, device="cuda")
    plan.run(din, dout, dl, ds)
    torch.cuda.synchronize()
    plan.set_timing(True)
    plan.run(din, dout, dl, ds)
    torch.cuda.synchronize()
    tm = plan.timings()
    tot = sum(tm.values())
    row = {"size": size, "buffers": nbuf, "ms": round(tot, 2), "gbs": round(size * nbuf / tot / 1e6, 2), "kernels_ms": {k: round(v, 2) for k, v in tm.items()}}
    rows.append(row)
    print(json.dumps(row), flush=True)
    plan.close()
    del din, dout
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "small_buffer_timing.json"), "w"), indent=1)
