"""checksum kernels timed inside a plan (device events), for the roofline table"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import sharpziplib_b200 as z
z.init(0)
n, sz = 1024, 262144
res = {}
for wrap, name in ((2, "crc32"), (1, "adler32")):
    plan = z.DeflatePlan([sz] * n, level=0, wrap=wrap)   # level 0: stored copy + checksum only
    din = torch.randint(0, 256, (plan.in_bytes,), dtype=torch.uint8, device="cuda")
    dout = torch.empty(plan.out_bytes, dtype=torch.uint8, device="cuda")
    dl = torch.zeros(n, dtype=torch.int64, device="cuda"); ds = torch.zeros(n, dtype=torch.int32, device="cuda")
    ck = torch.zeros(n, dtype=torch.int32, device="cuda")
    plan.set_timing(True)
    acc = {}
    for it in range(6):
        plan.run(din, dout, dl, ds, ck)
        torch.cuda.synchronize()
        if it >= 2:
            for k, v in plan.timings().items():
                acc[k] = acc.get(k, 0) + v / 4
    res[name] = {"ms": acc, "checksum_gbs": n * sz / acc["checksum"] / 1e6, "stored_copy_gbs": 2 * n * sz / acc["k_stored"] / 1e6}
print(json.dumps(res))
