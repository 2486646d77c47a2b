"""GPU tier, boxes with at least two GPUs: the static-table broadcast through NCCL inside the library
(b200z_static_tables_broadcast), one process per GPU under torchrun.  Skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_static_tables_broadcast_through_the_librarys_nccl_call(z):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    n = min(n, 4)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "tools", "nccl_broadcast_check.py")], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0 and r.stdout.count("ncclBroadcast ok") == n, (r.stdout[-2000:], r.stderr[-3000:])


def test_broadcast_rejects_a_missing_communicator(z):
    assert z.lib().b200z_static_tables_broadcast(None, 0, 0, None) == 1  # B200Z_E_ARG
