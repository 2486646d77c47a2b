"""CPU tier: data generator determinism, multi-rank sharding and the world_size-2 gloo path."""
import hashlib
import json
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "oracle_digests.json")


def test_datagen_is_deterministic_and_pinned():
    from sharpziplib_b200 import datagen
    pins = json.load(open(GOLD))["datagen_sha256_first_64k"]
    for i in range(8):
        d = datagen.silesia_mix(i, 65536)
        assert hashlib.sha256(d.tobytes()).hexdigest() == pins[datagen.CLASS_NAMES[i]]


def test_oracle_digests_pinned(oracle):
    # the oracle + generator produce the same compressed bytes on every box (self-consistency, not reference parity)
    from sharpziplib_b200 import datagen
    pins = json.load(open(GOLD))["oracle_deflate_l6_sha256_64k"]
    for i in range(8):
        c = oracle.deflate(datagen.silesia_mix(i, 65536).tobytes(), level=6)
        assert hashlib.sha256(c).hexdigest() == pins[datagen.CLASS_NAMES[i]]


def test_partition_by_bytes():
    from sharpziplib_b200.sharding import partition_by_bytes
    lens = [262144] * 1024
    for w in (1, 2, 4, 8):
        parts = partition_by_bytes(lens, w)
        assert parts[0][0] == 0 and parts[-1][1] == 1024
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        assert all(hi - lo == 1024 // w for lo, hi in parts)
    parts = partition_by_bytes([10, 1000, 10, 10, 1000, 10], 2)
    assert parts[0][1] == parts[1][0] and parts[1][1] == 6
    assert partition_by_bytes([], 4) == [(0, 0)] * 4


WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from sharpziplib_b200.sharding import partition_by_bytes, broadcast_static_tables
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
blob = broadcast_static_tables(dist)
lens = [1000 + 37 * i for i in range(101)]
lo, hi = partition_by_bytes(lens, dist.get_world_size())[dist.get_rank()]
mine = torch.tensor([sum(lens[lo:hi]), hi - lo], dtype=torch.int64)
allv = [torch.zeros(2, dtype=torch.int64) for _ in range(dist.get_world_size())]
dist.all_gather(allv, mine)
if dist.get_rank() == 0:
    assert sum(int(v[0]) for v in allv) == sum(lens) and sum(int(v[1]) for v in allv) == len(lens)
    print("GLOO_OK", len(blob))
dist.destroy_process_group()
'''


def test_world_size_2_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GLOO_OK" in outs[0]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line on stdout with the contract's keys"""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config"):
        assert k in d, k
