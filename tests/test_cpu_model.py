"""CPU tier: the kernel decomposition (b200z_core.cuh, the same __host__ __device__ code the CUDA kernels call) executed
serially by tests/cpu_model/model.cpp must reproduce the oracle bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import corpus_small

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    so = os.path.join(ROOT, "tests", "cpu_model", "_build", "libmodel.so")
    src = os.path.join(ROOT, "tests", "cpu_model", "model.cpp")
    hdr = os.path.join(ROOT, "sharpziplib_b200", "csrc", "b200z_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(so), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.dirname(hdr), "-o", so, src])
    M = C.CDLL(so)
    M.model_deflate.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]

    def run(d, level=6, strategy=0, flush_then_finish=0):
        a = np.frombuffer(d, dtype=np.uint8) if len(d) else np.zeros(1, np.uint8)
        cap = len(d) + len(d) // 8 + 1024
        out = np.zeros(cap, np.uint8)
        ol = C.c_uint64(0)
        rc = M.model_deflate(a.ctypes.data, len(d), level, strategy, flush_then_finish, out.ctypes.data, cap, C.byref(ol))
        assert rc == 0, rc
        return out[:ol.value].tobytes()
    return run


def test_model_matches_oracle(model, oracle):
    for name, d in corpus_small():
        for level in (5, 6, 9):
            assert model(d, level) == oracle.deflate(d, level=level), (name, level)
        assert model(d, 6, 0, 1) == oracle.deflate(d, level=6, pattern=1), name


def test_model_strategies(model, oracle):
    for name, d in corpus_small():
        if len(d) not in (1000, 70000):
            continue
        for strategy in (1, 2):
            assert model(d, 6, strategy) == oracle.deflate(d, level=6, strategy=strategy), (name, strategy)


def test_model_multi_slide(model, oracle):
    from sharpziplib_b200 import datagen
    for cls in (0, 4, 6):
        d = datagen.silesia_mix(cls, 400000, config=8).tobytes()
        assert model(d, 6) == oracle.deflate(d, level=6), cls


def test_model_fast_and_stored_levels(model, oracle):
    """levels 1-4 (serial DeflateFast engine, what k_fast runs) and level 0 (stored_run bookkeeping)"""
    from sharpziplib_b200 import datagen
    items = [(n, d) for n, d in corpus_small() if len(d) in (0, 3, 100, 4096, 70000)]
    items.append(("mix3_400k", datagen.silesia_mix(3, 400000, config=8).tobytes()))
    items.append(("zeros_300k", bytes(300000)))
    for name, d in items:
        for level in (0, 1, 2, 3, 4):
            assert model(d, level) == oracle.deflate(d, level=level), (name, level)
            assert model(d, level, 0, 1) == oracle.deflate(d, level=level, pattern=1), (name, level)
