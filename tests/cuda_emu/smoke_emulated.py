"""TEST INFRASTRUCTURE: a small end-to-end pass over every kernel family on the CUDA emulator, against the oracle.
Run by tests/test_cuda_emu.py in a subprocess (the emulator library must never sit in a process that also tests the real
one).  Exit code 0 = everything matched."""
import io
import os
import sys
import zlib

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_emulated  # noqa: E402

run_emulated.install(sanitize="asan" if "--asan" in sys.argv else "--sanitize" in sys.argv)
import numpy as np  # noqa: E402
import oracle_lib as O  # noqa: E402
import sharpziplib_b200 as z  # noqa: E402
from helpers import oracle_calls  # noqa: E402
from sharpziplib_b200 import datagen  # noqa: E402

z.init(0)
bufs = [datagen.silesia_mix(0, 20000).tobytes(), datagen.silesia_mix(3, 9000).tobytes(), b"", b"a", bytes(5000)]

# deflate: lazy levels (k_links, k_match, k_parse_*, k_plan, k_scan, k_emit), greedy levels (k_fast), stored (k_stored)
for level in (6, 9, 1, 0):
    outs, _ = z.deflate_batch(bufs, level=level)
    assert outs == [O.deflate(b, level=level) for b in bufs], level
# k_links with longer runs (B200Z_LINK_RUN): a stream of two 64 Ki runs / one 128 Ki run + a tail
long_buf = [datagen.silesia_mix(3, 140000).tobytes()]
long_ref = [O.deflate(long_buf[0], level=6)]
for run in ("65536", "131072"):
    os.environ["B200Z_LINK_RUN"] = run
    outs, _ = z.deflate_batch(long_buf, level=6)
    assert outs == long_ref, "B200Z_LINK_RUN=" + run
del os.environ["B200Z_LINK_RUN"]
# the parse: chunks entering from their warm-up state (default) and from the clean guess (k_parse_fix parses again at most boundaries)
os.environ["B200Z_PARSE_WARM"] = "0"
outs, _ = z.deflate_batch(long_buf, level=6)
assert outs == long_ref, "B200Z_PARSE_WARM=0"
del os.environ["B200Z_PARSE_WARM"]
# k_fast: group steps with head[] in the pool and in shared memory, the lane-0 statement, a window slide inside a stream
fast_bufs = [datagen.silesia_mix(1, 70000).tobytes(), datagen.gen_text(3000, 9).tobytes(), b"ab" * 2000]
for level in (2, 3, 4, 1):
    refs = [O.deflate(b, level=level) for b in fast_bufs]
    for head in (("pool", "smem") if level == 1 else ("pool",)):
        os.environ["B200Z_FAST_HEAD"] = head
        outs, _ = z.deflate_batch(fast_bufs, level=level)
        assert outs == refs, (level, head)
os.environ["B200Z_FAST_GROUP"] = "0"
outs, _ = z.deflate_batch(fast_bufs[1:], level=1)
assert outs == [O.deflate(b, level=1) for b in fast_bufs[1:]]
del os.environ["B200Z_FAST_GROUP"], os.environ["B200Z_FAST_HEAD"]
outs, checks = z.deflate_batch(bufs[:2], level=6, wrap=1)  # zlib framing: Adler-32 on the device (k_checksum)
assert outs == [O.deflate(b, level=6, nowrap=False) for b in bufs[:2]]

# inflate (block-parallel pipeline: k_find .. k_resolve, serial k_inflate for what it hands back), framed inflate (k_wrap_head / k_wrap_tail), truncation and restart points
comp = [O.deflate(b, level=6) for b in bufs]
back, used, st = z.inflate_batch(comp, [len(b) for b in bufs])
assert back == bufs and [int(u) for u in used] == [len(c) for c in comp]
zl = [zlib.compress(b, 6) for b in bufs[:2]]
back, used, st = z.inflate_batch(zl, [len(b) for b in bufs[:2]], wrap=1)
assert back == bufs[:2]
back, used, st = z.inflate_batch([comp[0][:3000]], [len(bufs[0])], raise_on_error=False)
assert int(st[0]) & 0xFF == 8 and bufs[0].startswith(back[0])

# checksums
for b in bufs:
    c = z.Crc32(); c.Update(b)
    a = z.Adler32(); a.Update(b)
    assert c.Value == zlib.crc32(b) and a.Value == zlib.adler32(b)

# entry ciphers (k_pbkdf2, k_aes_ctr, k_hmac_sha1, k_pkzip): a transform fed in pieces, a batch, both directions
from sharpziplib_b200 import encryption as E  # noqa: E402
for kb in (16, 32):
    salt, pw, data = bytes(range(kb // 2)), "pässword", datagen.silesia_mix(2, 5003).tobytes()
    want_ct, want_pv, want_mac = O.zip_aes(pw.encode(), salt, kb, True, data)
    t = E.ZipAESTransform(pw, salt, kb, True)
    out, pos = bytearray(len(data)), 0
    for piece in (1, 15, 16, 17, 700, 10 ** 6):
        k = min(piece, len(data) - pos)
        t.TransformBlock(data, pos, k, out, pos)
        pos += k
    assert bytes(out) == want_ct and t.PwdVerifier == want_pv and t.GetAuthCode() == want_mac, kb
    t.Dispose()
    keys = E.aes_derive_keys([pw.encode(), b""], [salt, salt], kb)
    back, auth = E.aes_batch([want_ct, b""], keys, kb, False)
    assert back[0] == data and auth[0].tobytes() == want_mac
k12 = E.PkzipClassic.GenerateKeys(b"secret")
assert k12 == O.pkzip_generate_keys(b"secret")
enc, after = E.pkzip_batch([bufs[0], bufs[4]], np.stack([np.frombuffer(k12, np.uint8)] * 2), True)
assert enc[0] == O.pkzip_transform(k12, True, bufs[0])[0] and after[0].tobytes() == O.pkzip_transform(k12, True, bufs[0])[1]
dec, _ = E.pkzip_batch(enc, np.stack([np.frombuffer(k12, np.uint8)] * 2), False)
assert dec == [bufs[0], bufs[4]]

# handles: a SetInput schedule with a Flush in the middle at a greedy level and at level 0; an Inflater fed in pieces
text = datagen.gen_text(30000, 4).tobytes()
for level in (1, 0, 6):
    segs = [(text[:12000], [5000, 7000]), (text[12000:], [18000])]
    d = z.Deflater(level, True)
    out = bytearray()
    buf = bytearray(1 << 16)
    for i, (seg, chunks) in enumerate(segs):
        pos = 0
        for c in chunks:
            d.SetInput(seg[pos:pos + c]); pos += c
            while True:
                k = d.Deflate(buf)
                if k <= 0:
                    break
                out += buf[:k]
        d.Flush() if i == 0 else d.Finish()
        while True:
            k = d.Deflate(buf)
            if k <= 0:
                break
            out += buf[:k]
    assert bytes(out) == oracle_calls(level, segs, None, True, nowrap=True), level
c = O.deflate(text, level=6, nowrap=False)
ins = z.InflaterInputStream(io.BytesIO(c), z.Inflater(False), 1024)
assert ins.read() == text
print("emulated smoke ok")
