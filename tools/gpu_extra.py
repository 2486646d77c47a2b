"""Extra measurements on the GPU box: checksum kernels vs the HBM roofline, and the C4 shape (one long stream)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import oracle_lib as O  # noqa: E402
import sharpziplib_b200 as z  # noqa: E402
from sharpziplib_b200 import datagen  # noqa: E402

z.init(0)
res = {}
# ---- checksums: 256 x 1 MiB resident on the device -------------------------------------------------------
n, sz = 256, 1 << 20
blob = torch.randint(0, 256, (n * sz,), dtype=torch.uint8, device="cuda")
off = (np.arange(n, dtype=np.int64) * sz)
lens = np.full(n, sz, dtype=np.int64)
for kind, name in ((0, "crc32"), (1, "adler32")):
    v = torch.zeros(n, dtype=torch.int32, device="cuda")
    def run():
        v.fill_(1 if kind else 0)
        rc = z.lib().b200z_checksum_batch_device(kind, blob.data_ptr(), off.ctypes.data, lens.ctypes.data, n, v.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    res[name + "_gbs_incl_host_setup"] = n * sz / ms / 1e6
    host = blob[:sz].cpu().numpy().tobytes()
    ref = O.crc32(host) if kind == 0 else O.adler32(host)
    assert int(v[0].item()) & 0xFFFFFFFF == ref, name
# ---- C4 shape: one long log stream, level 6 -----------------------------------------------------------------
size = int(os.environ.get("C4_MIB", "64")) << 20
d = datagen.log_stream(size).tobytes()
t = time.time()
out, _ = z.deflate_batch([d], level=6)
res["c4_deflate_s_host_api"] = time.time() - t
t = time.time()
ref = O.deflate(d, level=6)
res["c4_oracle_s"] = time.time() - t
res["c4_parity"] = out[0] == ref
res["c4_ratio"] = len(d) / len(ref)
t = time.time()
back, used, st = z.inflate_batch([ref], [len(d)])
res["c4_inflate_s_host_api"] = time.time() - t
res["c4_roundtrip"] = back[0] == d
plan = z.DeflatePlan([len(d)], level=6)
h = np.zeros(plan.in_bytes, dtype=np.uint8)
h[:len(d)] = np.frombuffer(d, dtype=np.uint8)
din = torch.from_numpy(h).cuda()
dout = torch.empty(plan.out_bytes, dtype=torch.uint8, device="cuda")
dl = torch.zeros(1, dtype=torch.int64, device="cuda")
ds = torch.zeros(1, dtype=torch.int32, device="cuda")
plan.set_timing(True)
plan.run(din, dout, dl, ds)
plan.run(din, dout, dl, ds)
torch.cuda.synchronize()
res["c4_kernels_ms"] = plan.timings()
print(json.dumps(res))
