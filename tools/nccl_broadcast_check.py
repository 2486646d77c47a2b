"""One rank per GPU under torchrun: creates an NCCL communicator the way a host program would (ncclGetUniqueId on rank 0, the id
shared out of band, ncclCommInitRank) and hands it to b200z_static_tables_broadcast -- the library's own ncclBroadcast of the
static Huffman tables (north_star's only collective), no torch.distributed on that path.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/nccl_broadcast_check.py"""
import ctypes as C
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import sharpziplib_b200 as z  # noqa: E402
from sharpziplib_b200 import _lib  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dry = os.environ.get("B200Z_NCCL_DRY") == "1"  # no GPU: everything up to the communicator (checks the plumbing of this script)
if not dry:
    torch.cuda.set_device(local)
    z.init(local)
dist.init_process_group("gloo")  # only carries the 128-byte id: the out-of-band channel every NCCL program needs
cands = [f for p in sys.path for f in glob.glob(os.path.join(p, "nvidia", "nccl", "lib", "libnccl.so*"))] + ["libnccl.so.2"]
N = C.CDLL(cands[0], mode=C.RTLD_GLOBAL)  # the NCCL torch ships; the library finds the same one (dlopen ... RTLD_NOLOAD)


class UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


uid = UniqueId()
if rank == 0:
    assert N.ncclGetUniqueId(C.byref(uid)) == 0
t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
dist.broadcast(t, src=0)
C.memmove(C.byref(uid), t.numpy().tobytes(), 128)
if dry:
    print("dry run: id %s..." % bytes(uid)[:8].hex(), cands[0])
    dist.destroy_process_group()
    sys.exit(0)
comm = C.c_void_p()
N.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
assert N.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
rc = _lib.lib().b200z_static_tables_broadcast(comm, 0, rank, None)
_lib.raise_for(rc)
N.ncclCommDestroy.argtypes = [C.c_void_p]
N.ncclCommDestroy(comm)
dist.barrier()
print("rank %d of %d: static tables broadcast over the library's ncclBroadcast ok" % (rank, world), flush=True)
dist.destroy_process_group()
