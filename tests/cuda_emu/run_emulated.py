"""TEST INFRASTRUCTURE: runs GPU-tier tests against the CUDA emulator (tests/cuda_emu) in THIS process only.

    python tests/cuda_emu/run_emulated.py [pytest args ...]     e.g.  -k "deflate_parity and 6"  tests/test_gpu_parity.py

The product package is not touched: the emulator library is put where sharpziplib_b200._lib keeps its ctypes handle before
any test runs, and torch's device entry points the tests use are mapped onto host tensors ("device memory" of the emulator
is host memory).  Slow (every CUDA thread is a fiber); meant for checking kernels on small inputs when there is no GPU."""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)


def install(sanitize=False):
    import build_emu
    so = build_emu.build(sanitize=sanitize)
    from sharpziplib_b200 import _lib
    L = C.CDLL(so)
    for name, (res, args) in _lib._SIGS.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib._lib = L
    # torch: the tests build device tensors for the plan API; under the emulator they are host tensors
    import torch
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None

    class _S:
        cuda_stream = 0
    torch.cuda.current_stream = lambda *a, **k: _S()
    torch.Tensor.is_cuda = property(lambda self: True)
    for fn in ("zeros", "empty", "ones"):
        orig = getattr(torch, fn)

        def wrap(*a, _orig=orig, **k):
            k.pop("device", None)
            return _orig(*a, **k)
        setattr(torch, fn, wrap)
    return so


if __name__ == "__main__":
    import pytest
    args = sys.argv[1:]
    # --sanitize: UBSan build; --asan: AddressSanitizer build (start python with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
    # ASAN_OPTIONS=detect_leaks=0): "device" buffers are heap memory, out-of-bounds kernel accesses get reported
    sanitize = "asan" if "--asan" in args else "--sanitize" in args
    args = [a for a in args if a not in ("--sanitize", "--asan")]
    print("emulator library:", install(sanitize))
    sys.exit(pytest.main(["-q", "-m", "gpu", "-p", "no:cacheprovider"] + args))
