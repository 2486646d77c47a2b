"""TEST INFRASTRUCTURE: sizes around the tile / run / window / round boundaries of the kernels, on the CUDA emulator (best
under --asan: "device" buffers are heap memory, so reads and writes outside them are reported).
    LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
        python tests/cuda_emu/boundary_sweep.py --asan"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_emulated  # noqa: E402

run_emulated.install(sanitize="asan" if "--asan" in sys.argv else "--sanitize" in sys.argv)
import oracle_lib as O  # noqa: E402
import sharpziplib_b200 as z  # noqa: E402
from sharpziplib_b200 import datagen  # noqa: E402

z.init(0)
sizes = [1, 2, 3, 4, 15, 16, 17, 31, 32, 33, 1023, 1024, 1025, 16383, 16384, 16385, 32505, 32506, 32507, 32767, 32768, 32769,
         65273, 65274, 65275, 65535, 65536, 65537, 98303, 98304, 98305]
if "--quick" in sys.argv:
    sizes = [s for s in sizes if s < 40000]
src = {0: datagen.silesia_mix(0, 100000).tobytes(), 4: datagen.silesia_mix(4, 100000).tobytes(), 7: datagen.silesia_mix(7, 100000).tobytes()}
t0 = time.time()
for level in (6, 1, 0):
    for n in sizes:
        bufs = [src[c][:n] for c in (0, 4, 7)]
        outs, _ = z.deflate_batch(bufs, level=level)
        assert outs == [O.deflate(b, level=level) for b in bufs], (level, n)
        if level == 6:
            back, used, st = z.inflate_batch(outs, [len(b) for b in bufs])
            assert back == bufs, n
            # the output capacity exactly met, and one byte short
            back, used, st = z.inflate_batch(outs[:1], [n - 1], raise_on_error=False)
            assert int(st[0]) & 0xFF == 7, (n, st)
            c = z.Crc32(); c.Update(bufs[0])
            assert c.Value == O.crc32(bufs[0])
    print("level", level, "ok  %.0f s" % (time.time() - t0), flush=True)
print("boundary sweep ok")
