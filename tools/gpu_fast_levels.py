"""levels 1-4 throughput by buffer size: k_fast with its warp-wide group steps against the same kernel with every loop top on
lane 0 (B200Z_FAST_GROUP=0), parity of several buffers per point against the oracle.
    python tools/gpu_fast_levels.py [levels, e.g. 1,3] [--big]"""
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
levels = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 and sys.argv[1][0] != "-" else [1, 2, 3, 4]
shapes = [(4096, 16384), (65536, 2048), (262144, 1024)]
if "--big" in sys.argv:
    shapes.append((64 << 20, 4))
rows = []
for level in levels:
    for size, nbuf in shapes:
        ndist = min(nbuf, 32 if size < (1 << 20) else 4)
        bufs = [datagen.silesia_mix(i, size, config=5) for i in range(ndist)]
        plan = z.DeflatePlan([size] * nbuf, level=level)
        h = np.zeros(plan.in_bytes, np.uint8)
        for i, o in enumerate(plan.in_offsets):
            h[o:o + size] = bufs[i % ndist]
        din = torch.from_numpy(h).cuda()
        dout = torch.empty(plan.out_bytes, dtype=torch.uint8, device="cuda")
        dl = torch.zeros(nbuf, dtype=torch.int64, device="cuda")
        ds = torch.zeros(nbuf, dtype=torch.int32, device="cuda")
        row = {"level": level, "size": size, "buffers": nbuf}
        for mode in ("1", "0"):
            if mode == "0" and size >= (1 << 24):
                continue  # the serial statement takes seconds per 64 MiB stream
            os.environ["B200Z_FAST_GROUP"] = mode
            plan.run(din, dout, dl, ds)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run(din, dout, dl, ds)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            plan.set_timing(True)
            plan.run(din, dout, dl, ds)
            torch.cuda.synchronize()
            tm = plan.timings()
            plan.set_timing(False)
            ok = True
            for i in range(min(ndist, 8)):
                o = int(plan.out_offsets[i])
                ok &= dout[o:o + int(dl[i])].cpu().numpy().tobytes() == O.deflate(bufs[i].tobytes(), level=level)
            row["group" if mode == "1" else "serial"] = {"ms": round(ms, 3), "gbs": round(size * nbuf / ms / 1e6, 3), "parity": bool(ok),
                                                              "kernels_ms": {k: round(float(v), 3) for k, v in tm.items()}}
        os.environ["B200Z_FAST_GROUP"] = "1"
        rows.append(row)
        print(json.dumps(row), flush=True)
        del plan, din, dout
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "fast_levels.json"), "w"), indent=1)
