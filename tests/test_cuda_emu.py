"""CPU tier: the kernel SOURCES run on the CUDA emulator (tests/cuda_emu: every CUDA thread a fiber) against the oracle -- one
small pass over every kernel family, so that a kernel edit that breaks bit-exactness shows up without a GPU.  In a
subprocess: the emulator library never shares a process with tests of the real one."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kernels_on_the_emulator_match_the_oracle():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cuda_emu", "smoke_emulated.py")], capture_output=True,
                       text=True, timeout=1500)
    assert r.returncode == 0 and "emulated smoke ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


def test_emulator_schedule_orders_expose_a_race(tmp_path):
    """B200Z_EMU_ORDER self-test: a kernel that reads its neighbour's shared-memory slot without a barrier gives
    order-dependent results, the same kernel with the barrier does not."""
    emu = os.path.join(ROOT, "tests", "cuda_emu")
    exe = str(tmp_path / "order_selftest")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-w", "-I", emu, "-I", os.path.join(emu, "include"),
                           os.path.join(emu, "order_selftest.cpp"), os.path.join(emu, "cuda_emu.cpp"), "-o", exe])
    res = {}
    for order in ("fwd", "rev", "rand:3"):
        out = subprocess.run([exe], env=dict(os.environ, B200Z_EMU_ORDER=order), capture_output=True, text=True, timeout=60).stdout
        res[order] = dict(kv.split("=") for kv in out.split())
    assert len({r["clean"] for r in res.values()}) == 1
    assert len({r["racy"] for r in res.values()}) == 3


def test_kernels_are_schedule_independent_on_the_emulator():
    """the same pass with the threads of every block resumed in descending and in pseudo-random order (and the blocks of
    every grid likewise): a result that depended on it would be a race the default order hides"""
    for order in ("rev", "rand:11"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "cuda_emu", "smoke_emulated.py")], capture_output=True,
                           text=True, timeout=1500, env=dict(os.environ, B200Z_EMU_ORDER=order))
        assert r.returncode == 0 and "emulated smoke ok" in r.stdout, (order, r.stdout[-2000:], r.stderr[-4000:])
