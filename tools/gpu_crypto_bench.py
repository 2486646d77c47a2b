"""Entry ciphers on device memory (b200z_aes_device / b200z_pkzip_device): throughput on the compressed bytes of a C3-like batch
(1024 entries x 128 KiB) and on one 256 MiB entry, CUDA events, results checked against the oracle on a few entries."""
import ctypes as C
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
from sharpziplib_b200 import _lib, encryption as E  # noqa: E402

z.init(0)
L = _lib.lib()
rows = []
for n, size, kb in ((1024, 128 << 10, 32), (1024, 128 << 10, 16), (16384, 4096, 32), (1, 256 << 20, 32)):
    rng = np.random.default_rng(n)
    stride = (size + 255) // 256 * 256
    h = rng.integers(0, 256, n * stride, dtype=np.uint8)
    d_in = torch.from_numpy(h).cuda()
    d_out = torch.empty_like(d_in)
    off = torch.arange(n, dtype=torch.int64) * stride
    ln = torch.full((n,), size, dtype=torch.int64)
    d_off, d_len = off.cuda(), ln.cuda()
    pws = [b"password%d" % i for i in range(n)]
    salts = [rng.bytes(kb // 2) for _ in range(n)]
    keys = E.aes_derive_keys(pws, salts, kb)
    d_keys = torch.from_numpy(keys.reshape(-1).copy()).cuda()
    d_state = torch.zeros(n * int(L.b200z_aes_state_bytes()), dtype=torch.uint8, device="cuda")
    d_auth = torch.zeros(n * 20, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream

    def run():
        d_state.zero_()
        _lib.raise_for(L.b200z_aes_device(d_in.data_ptr(), d_out.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), size, n, kb,
                                          d_keys.data_ptr(), 1, d_state.data_ptr(), 1, d_auth.data_ptr(), s))
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ok = True
    for i in ([0, n // 2, n - 1] if size <= (1 << 20) else [0]):
        ct, _, mac = O.zip_aes(pws[i], salts[i], kb, True, h[i * stride:i * stride + size].tobytes())
        ok &= d_out[i * stride:i * stride + size].cpu().numpy().tobytes() == ct and d_auth[20 * i:20 * i + 20].cpu().numpy().tobytes() == mac
    row = {"cipher": "aes%d-ctr+hmac-sha1" % (8 * kb), "entries": n, "size": size, "ms": round(ms, 3), "gbs": round(n * size / ms / 1e6, 2), "parity": bool(ok)}
    rows.append(row)
    print(json.dumps(row), flush=True)
    if size <= (1 << 20):
        k12 = np.stack([np.frombuffer(O.pkzip_generate_keys(p), np.uint8) for p in pws]).view(np.uint32).reshape(-1).copy()
        d_k = torch.from_numpy(k12.view(np.int32)).cuda()
        for enc in (1,):
            d_k.copy_(torch.from_numpy(k12.view(np.int32)))
            torch.cuda.synchronize()
            e0.record()
            _lib.raise_for(L.b200z_pkzip_device(d_in.data_ptr(), d_out.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), n, d_k.data_ptr(), enc, s))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            w, _ = O.pkzip_transform(O.pkzip_generate_keys(pws[1]), True, h[stride:stride + size].tobytes())
            ok = d_out[stride:stride + size].cpu().numpy().tobytes() == w
            row = {"cipher": "pkzip-classic", "entries": n, "size": size, "ms": round(ms, 3), "gbs": round(n * size / ms / 1e6, 2), "parity": bool(ok)}
            rows.append(row)
            print(json.dumps(row), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "crypto_bench.json"), "w"), indent=1)
