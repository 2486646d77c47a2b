"""Multi-GPU layout of a batch: independent streams are split across ranks by byte count; the only collective on the
path is the one-time broadcast of the static Huffman tables (north_star / SURVEY.md 8e).  One process per GPU."""
import numpy as np

from . import _lib


def partition_by_bytes(lens, world_size):
    """Contiguous index ranges [lo, hi) per rank, balanced by cumulative byte count: rank r gets the buffers whose
    cumulative start offset falls in [r*T/R, (r+1)*T/R)."""
    lens = np.ascontiguousarray(lens, dtype=np.int64)
    first = np.zeros(world_size + 1, dtype=np.int32)
    # the library's own cut (b200z_partition_by_bytes): what b200z_*_batch_multi uses for one process and several GPUs
    _lib.raise_for(_lib.lib().b200z_partition_by_bytes(lens.ctypes.data, lens.size, world_size, first.ctypes.data))
    return [(int(first[r]), int(first[r + 1])) for r in range(world_size)]


def broadcast_static_tables(dist, device=None):
    """Rank 0 exports the static Huffman tables, every rank receives them through torch.distributed (NCCL on GPUs,
    gloo in the CPU tests) and installs/verifies them.  Returns the blob."""
    import ctypes as C

    import torch
    L = _lib.lib()
    n = L.b200z_static_tables_size()
    buf = (C.c_uint8 * n)()
    if dist.get_rank() == 0:
        _lib.raise_for(L.b200z_static_tables_export(buf, n))
    t = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    blob = bytes(t.cpu().numpy().tobytes())
    arr = (C.c_uint8 * n).from_buffer_copy(blob)
    _lib.raise_for(L.b200z_static_tables_import(arr, n))
    return blob
