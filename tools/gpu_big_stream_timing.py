"""per-kernel times of level-6 deflate plans over few large streams (what bounds C4 and the large end of C5)"""
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
for name, sizes in (("4x64MiB mix", [64 << 20] * 4), ("1x512MiB log", [512 << 20]), ("1024x256KiB mix", [256 << 10] * 1024)):
    if "log" in name:
        bufs = [datagen.log_stream(sizes[0], config=4)]
    else:
        base = [datagen.silesia_mix(i, sizes[0], config=5) for i in range(min(len(sizes), 8))]
        bufs = [base[i % len(base)] for i in range(len(sizes))]
    plan = z.DeflatePlan(sizes, level=6)
    h = np.zeros(plan.in_bytes, np.uint8)
    for i, o in enumerate(plan.in_offsets):
        h[o:o + sizes[i]] = bufs[i]
    din = torch.from_numpy(h).cuda()
    dout = torch.empty(plan.out_bytes, dtype=torch.uint8, device="cuda")
    dl = torch.zeros(len(sizes), dtype=torch.int64, device="cuda")
    ds = torch.zeros(len(sizes), dtype=torch.int32, device="cuda")
    plan.run(din, dout, dl, ds)
    torch.cuda.synchronize()
    plan.set_timing(True)
    plan.run(din, dout, dl, ds)
    torch.cuda.synchronize()
    tm = plan.timings()
    tot = sum(tm.values())
    row = {"workload": name, "ms": round(tot, 2), "gbs": round(sum(sizes) / tot / 1e6, 2), "kernels_ms": {k: round(v, 2) for k, v in tm.items()}}
    rows.append(row)
    print(json.dumps(row), flush=True)
    plan.close()
    del din, dout
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "big_stream_timing.json"), "w"), indent=1)
