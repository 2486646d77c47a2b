"""Small fixed workload for ncu captures: one deflate plan run (ND x 256 KiB Silesia-mix, level 6) and one inflate plan
run (NI x 1 MiB text) after a warm-up run of each."""
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
nd, ni = int(os.environ.get("ND", "128")), int(os.environ.get("NI", "32"))
d_np = [datagen.silesia_mix(i, 262144) for i in range(nd)]
t_np = [datagen.text_buffer(i, 1 << 20) for i in range(ni)]
comp = O.batch(0, [a.tobytes() for a in t_np], level=6, threads=16)
dplan = z.DeflatePlan([a.size for a in d_np], level=6)
iplan = z.InflatePlan([len(c) for c in comp], [a.size for a in t_np])
h = np.zeros(dplan.in_bytes, dtype=np.uint8)
for o, a in zip(dplan.in_offsets, d_np):
    h[o:o + a.size] = a
d_in = torch.from_numpy(h).cuda()
h2 = np.zeros(iplan.in_bytes, dtype=np.uint8)
for o, c in zip(iplan.in_offsets, comp):
    h2[o:o + len(c)] = np.frombuffer(c, dtype=np.uint8)
i_in = torch.from_numpy(h2).cuda()
d_out = torch.empty(dplan.out_bytes, dtype=torch.uint8, device="cuda")
i_out = torch.empty(iplan.out_bytes, dtype=torch.uint8, device="cuda")
dl = torch.zeros(nd, dtype=torch.int64, device="cuda")
ds = torch.zeros(nd, dtype=torch.int32, device="cuda")
il = torch.zeros(ni, dtype=torch.int64, device="cuda")
ist = torch.zeros(ni, dtype=torch.int32, device="cuda")
for _ in range(2):
    dplan.run(d_in, d_out, dl, ds)
    iplan.run(i_in, i_out, il, ist)
    torch.cuda.synchronize()
print("status", int(ds.abs().sum()), int(ist.abs().sum()), "clen", int(dl.sum()), "ilen", int(il.sum()))
