"""TEST INFRASTRUCTURE helper: levels 1-4 (k_fast with its warp-wide group steps) on the CUDA emulator (tests/cuda_emu) over a few
buffers of the generators, byte for byte against the oracle; with B200Z_EMU_CXXFLAGS=-DB200Z_FAST_STATS the kernel prints how
many positions a group step takes on average.
    B200Z_EMU_CXXFLAGS=-DB200Z_FAST_STATS python tools/emu_fast_check.py [size_kib] [levels, e.g. 1,3]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "cuda_emu"))
import run_emulated  # noqa: E402

run_emulated.install()
import oracle_lib as O  # noqa: E402
import sharpziplib_b200 as z  # noqa: E402
from sharpziplib_b200 import datagen  # noqa: E402

kib = int(sys.argv[1]) if len(sys.argv) > 1 else 96
levels = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3, 4]
z.init(0)
bufs = [datagen.silesia_mix(i, kib * 1024 - 13 * i, config=3).tobytes() for i in range(8)]
bufs.append(datagen.text_buffer(3, kib * 1024, config=2).tobytes())
bufs.append(datagen.gen_log(kib * 1024, 5).tobytes())
ok = True
for level in levels:
    t = time.time()
    got, _ = z.deflate_batch(bufs, level=level)
    for i, b in enumerate(bufs):
        want = O.deflate(b, level=level)
        good = bytes(got[i]) == want
        ok &= good
        print("level", level, "buffer", i, len(b), "->", len(want), "OK" if good else "MISMATCH")
    print("level %d: %.1f s" % (level, time.time() - t))
sys.exit(0 if ok else 1)
