"""First-contact diagnostic run on a GPU box: exercises every kernel on small inputs and reports mismatches in
detail instead of stopping at the first one.  (Development aid; the real tests are tests/ -m gpu.)"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle_lib as O  # noqa: E402
import sharpziplib_b200 as z  # noqa: E402
from sharpziplib_b200 import datagen  # noqa: E402


def section(name):
    print("=" * 20, name, flush=True)


def main():
    z.init(0)
    section("checksums")
    for n in [0, 1, 9, 127, 128, 129, 4096, 32768, 32769, 100000, 1 << 20]:
        d = datagen.gen_entropy(n, 77 + n).tobytes() if n else b""
        try:
            c = z.Crc32(); c.Update(d)
            a = z.Adler32(); a.Update(d)
            ok = (c.Value == O.crc32(d), a.Value == O.adler32(d))
            print(n, "crc", hex(c.Value), hex(O.crc32(d)), "adler", hex(a.Value), hex(O.adler32(d)), ok, flush=True)
        except Exception:
            traceback.print_exc()
    c = z.Crc32(); c.Update(b"123456789"); print("crc KAT", hex(c.Value))
    section("deflate")
    bad = 0
    for n in [0, 1, 5, 100, 4096, 40000, 70000, 262144, 300001]:
        for cls in range(8):
            d = datagen.silesia_mix(cls, n).tobytes() if n else b""
            for lvl in (6,) if cls else (5, 6, 7, 8, 9):
                try:
                    t = time.time()
                    out, _ = z.deflate_batch([d], level=lvl)
                    dt = time.time() - t
                    ref = O.deflate(d, level=lvl)
                    if out[0] != ref:
                        bad += 1
                        k = next((i for i in range(min(len(ref), len(out[0]))) if ref[i] != out[0][i]), -1)
                        print("MISMATCH n=%d cls=%d lvl=%d len ref=%d got=%d firstdiff=%d" % (n, cls, lvl, len(ref), len(out[0]), k), flush=True)
                    elif cls == 0:
                        print("ok n=%d lvl=%d clen=%d %.3fs" % (n, lvl, len(ref), dt), flush=True)
                except Exception:
                    bad += 1
                    traceback.print_exc()
        if n == 0:
            continue
    print("deflate mismatches:", bad, flush=True)
    section("deflate batch 64 x 256 KiB")
    bufs = [datagen.silesia_mix(i, 262144).tobytes() for i in range(64)]
    t = time.time()
    outs, _ = z.deflate_batch(bufs, level=6)
    print("batch time %.3fs" % (time.time() - t))
    refs = O.batch(0, bufs, level=6, threads=8)
    print("batch parity:", sum(1 for a, b in zip(outs, refs) if a == b), "/", len(bufs), flush=True)
    section("inflate")
    back, used, status = z.inflate_batch(refs, [len(b) for b in bufs], raise_on_error=False)
    print("inflate status", set(status.tolist()), "ok", sum(1 for a, b in zip(back, bufs) if a == b), "/", len(bufs))
    print("in_used ok", all(int(u) == len(r) for u, r in zip(used, refs)), flush=True)
    fix = bytes.fromhex("2b492d2e49cbcc495548cecf2b49cd2b29e60200")
    print("fixture", z.inflate_batch([fix], [100], raise_on_error=False))
    import zlib
    for lvl in (0, 1, 9):
        co = zlib.compressobj(lvl, zlib.DEFLATED, -15)
        zs = co.compress(bufs[0]) + co.flush()
        b2, u2, s2 = z.inflate_batch([zs], [len(bufs[0])], raise_on_error=False)
        print("zlib level", lvl, "status", s2, "ok", b2[0] == bufs[0], "used", int(u2[0]), len(zs), flush=True)
    section("handles / streams")
    try:
        import io
        for nowrap in (True, False):
            d = bufs[0][:100000]
            ms = io.BytesIO()
            ds = z.DeflaterOutputStream(ms, z.Deflater(6, nowrap))
            ds.IsStreamOwner = False
            ds.Write(d); ds.Flush(); ds.Finish()
            ref = O.deflate(d, level=6, nowrap=nowrap, pattern=1)
            print("stream flush/finish nowrap", nowrap, ms.getvalue() == ref, len(ref), len(ms.getvalue()), flush=True)
            ins = z.InflaterInputStream(io.BytesIO(ref), z.Inflater(nowrap))
            print("inflater stream", ins.read() == d, flush=True)
        ms = io.BytesIO()
        g = z.GZipOutputStream(ms); g.IsStreamOwner = False; g.ModifiedTime = 1577836800
        g.Write(bufs[2]); g.Finish()
        import gzip
        print("gzip ok", gzip.decompress(ms.getvalue()) == bufs[2], flush=True)
        print("gzip read", z.GZipInputStream(io.BytesIO(ms.getvalue())).read() == bufs[2], flush=True)
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
