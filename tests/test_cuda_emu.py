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
