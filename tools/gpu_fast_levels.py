"""levels 1-4 throughput by buffer size (k_fast runs one serial engine per stream; how many share an SM depends on prev[])"""
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
for level in (1, 3):
    for size, nbuf in ((4096, 16384), (16384, 4096), (65536, 1024), (262144, 1024)):
        bufs = [datagen.silesia_mix(i % 8, size, config=5) for i in range(min(nbuf, 64))]
        plan = z.DeflatePlan([size] * nbuf, level=level)
        h = np.zeros(plan.in_bytes, np.uint8)
        for i, o in enumerate(plan.in_offsets):
            h[o:o + size] = bufs[i % len(bufs)]
        din = torch.from_numpy(h).cuda()
        dout = torch.empty(plan.out_bytes, dtype=torch.uint8, device="cuda")
        dl = torch.zeros(nbuf, dtype=torch.int64, device="cuda")
        ds = torch.zeros(nbuf, dtype=torch.int32, device="cuda")
        plan.run(din, dout, dl, ds)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run(din, dout, dl, ds)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        o = int(plan.out_offsets[3])
        ok = dout[o:o + int(dl[3])].cpu().numpy().tobytes() == O.deflate(bufs[3 % len(bufs)].tobytes(), level=level)
        print("level %d  %7d B x %5d  %8.2f ms  %6.2f GB/s  parity %s" % (level, size, nbuf, ms, size * nbuf / ms / 1e6, ok), flush=True)
