"""TEST INFRASTRUCTURE helper: runs the block-parallel inflate pipeline on the CUDA emulator (tests/cuda_emu) over a few
multi-block streams and prints the pipeline's statistics (segments found, joins, passes).
    python tools/emu_inflate_check.py [size_kib] [n_streams]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
import run_emulated  # noqa: E402

run_emulated.install()
import numpy as np  # noqa: E402
import torch  # noqa: E402
import oracle_lib as O  # noqa: E402
import sharpziplib_b200 as z  # noqa: E402
from sharpziplib_b200 import datagen  # noqa: E402

kib = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
z.init(0)
orig = [datagen.text_buffer(i, kib * 1024 - 37 * i, config=2).tobytes() for i in range(n)]
orig += [datagen.silesia_mix(i, kib * 512 + 11 * i, config=3).tobytes() for i in range(n)]
comp = [O.deflate(b, level=6 if i % 2 == 0 else 9) for i, b in enumerate(orig)]
plan = z.InflatePlan([len(c) for c in comp], [len(b) + 64 for b in orig])
h_in = torch.zeros(plan.in_bytes, dtype=torch.uint8)
for o, c in zip(plan.in_offsets, comp):
    h_in[o:o + len(c)] = torch.frombuffer(bytearray(c), dtype=torch.uint8)
d_out = torch.zeros(plan.out_bytes, dtype=torch.uint8)
m = len(comp)
d_len = torch.zeros(m, dtype=torch.int64)
d_st = torch.zeros(m, dtype=torch.int32)
d_used = torch.zeros(m, dtype=torch.int64)
t = time.time()
plan.run(h_in, d_out, d_len, d_st, None, d_used)
print("ran in %.1f s" % (time.time() - t), plan.stats())
ok = True
for i, b in enumerate(orig):
    got = d_out[plan.out_offsets[i]:plan.out_offsets[i] + int(d_len[i])].numpy().tobytes()
    good = got == b and int(d_st[i]) == 0 and int(d_used[i]) == len(comp[i])
    ok &= good
    print(i, len(b), len(comp[i]), "status", int(d_st[i]), "out", int(d_len[i]), "used", int(d_used[i]), "OK" if good else "MISMATCH")
sys.exit(0 if ok else 1)
