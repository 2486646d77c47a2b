"""CPU tier: the oracle against the reference's own golden vectors and structural checks."""
import json
import os
import zlib

import numpy as np
import pytest

from helpers import corpus_small, crafted_t8

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.json")))


def test_checksum_kats(oracle):
    for k in GOLD["checksum_kats"]["crc32"]:
        assert oracle.crc32(k["ascii"].encode()) == k["value"]
    for k in GOLD["checksum_kats"]["adler32"]:
        assert oracle.adler32(k["ascii"].encode()) == k["value"]
    assert oracle.crc32(b"") == 0 and oracle.adler32(b"") == 1  # Reset values (ChecksumTests.cs:24-37)


def test_checksum_vs_zlib_and_bytewise(oracle):
    rng = np.random.default_rng(1)
    for n in (1, 15, 16, 17, 3799, 3800, 3801, 100000):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.crc32(d) == zlib.crc32(d)
        assert oracle.adler32(d) == zlib.adler32(d)
        L = oracle.lib()
        a = np.frombuffer(d, dtype=np.uint8)
        assert L.szl_crc32_bytewise(a.ctypes.data, n) == zlib.crc32(d)
        assert L.szl_adler32_bytewise(a.ctypes.data, n) == zlib.adler32(d)
        # running updates
        assert oracle.crc32_update(oracle.crc32(d[:n // 2]), d[n // 2:]) == zlib.crc32(d)
        assert oracle.adler32_update(oracle.adler32(d[:n // 2]), d[n // 2:]) == zlib.adler32(d)


def test_adler32_256mib_dotnet_random_kat(oracle):
    """ChecksumTests.Adler_32_Performance (test/.../Checksum/ChecksumTests.cs:41-64): Adler-32 of 256 MiB from
    `new Random(1).NextBytes`, then of "123456789", must be 0xD4897DA3.  Pins both the oracle's Adler32 on a long
    input (deferred modulo, Adler32.cs:134-161) and the restated System.Random the other reference-shaped tests use."""
    buf = oracle.dotnet_random_bytes(1, 256 * 1024 * 1024)
    a = oracle.adler32(buf)
    assert oracle.adler32_update(a, b"123456789") == 0xD4897DA3


@pytest.mark.parametrize("nowrap", [False, True])
def test_inflate_deflate_reference_shape(oracle, nowrap):
    """InflaterDeflaterTests.InflateDeflateZlib / NonZlib (test/.../Base/InflaterDeflaterTests.cs:157-162, :226-231)
    on the reference's own input, Utils.GetDummyBytes(100000) = Random(5): Write(all) -> Flush() -> Finish()."""
    original = oracle.dotnet_random_bytes(5, 100000).tobytes()
    for level in range(10):
        comp = oracle.deflate(original, level=level, nowrap=nowrap, pattern=1)
        assert zlib.decompress(comp, -15 if nowrap else 15) == original
        back, rem, fin = oracle.inflate(comp, nowrap=nowrap, max_out=len(original) + 64)
        assert back == original and fin and rem == 0


def test_inflater_reference_fixture(oracle):
    raw = bytes.fromhex(GOLD["inflate_ok"]["raw_hex"])
    out, remaining, finished = oracle.inflate(raw)
    assert out == b"testfile contents\n" and finished and remaining == 0
    assert zlib.crc32(out) == GOLD["inflate_ok"]["entry"]["crc"]


def test_inflater_zero_codelength_fixture(oracle):
    raw = bytes.fromhex(GOLD["inflate_zero_codelength"]["raw_hex"])
    with pytest.raises(oracle.OracleError) as e:
        oracle.inflate(raw, max_out=1 << 16)
    assert e.value.kind == 3 and "invalid codelength 0" in e.value.msg


def test_deflater_outputs_are_valid_deflate(oracle):
    for name, d in corpus_small():
        for level in range(10):
            for nowrap in (True, False):
                c = oracle.deflate(d, level=level, nowrap=nowrap)
                assert zlib.decompress(c, -15 if nowrap else 15) == d, (name, level)
                back, rem, fin = oracle.inflate(c, nowrap=nowrap, max_out=len(d) + 64)
                assert back == d and fin and rem == 0, (name, level)


def test_deflater_feed_pattern_invariance_lazy_levels(oracle):
    # trap T9: for levels 5-9 without a mid-stream Flush the bytes do not depend on how input is chunked
    for name, d in corpus_small():
        if len(d) < 1000:
            continue
        for level in (5, 6, 9):
            whole = oracle.deflate(d, level=level)
            for chunk in (1, 4096, 7777):
                if chunk == 1 and len(d) > 5000:
                    continue
                assert oracle.deflate(d, level=level, chunk=chunk) == whole, (name, level, chunk)
            assert oracle.deflate(d, level=level, pattern=2, chunk=4096, obuf=512) == whole


def test_deflater_known_structures(oracle):
    assert oracle.deflate(b"", 6) == bytes.fromhex("0300")                      # final static block with only EOB
    assert oracle.deflate(b"", 6, nowrap=False) == bytes.fromhex("789c030000000001")
    assert oracle.deflate(b"", 0, nowrap=False)[:2] == bytes.fromhex("78da")    # trap T11: level 0 header is 78 DA
    assert oracle.deflate(b"", 1, nowrap=False)[:2] == bytes.fromhex("7801")
    assert oracle.deflate(b"", 3, nowrap=False)[:2] == bytes.fromhex("785e")
    assert oracle.deflate(b"", 9, nowrap=False)[:2] == bytes.fromhex("78da")
    # Write -> Flush -> Finish on empty input: empty non-final static block, sync padding, final empty block (T6)
    assert oracle.deflate(b"", 6, pattern=1) == bytes.fromhex("020820c000")


def test_deflater_block_structure(oracle):
    from sharpziplib_b200 import datagen
    d = datagen.gen_text(300000, 1).tobytes()
    c, rows = oracle.deflate_trace(d, level=6)
    assert all(r[1] == 16384 for r in rows[:-1])          # blocks are cut every 16384 symbols only (trap T5)
    assert rows[-1][1] <= 16384
    rnd = datagen.Rng(2).bytes(100000).tobytes()
    c, rows = oracle.deflate_trace(rnd, level=6)
    assert any(r[0] == 0 for r in rows)                   # incompressible data goes to stored blocks


def test_t8_slide_sentinel(oracle):
    d = crafted_t8()
    c = oracle.deflate(d, level=6)
    assert zlib.decompress(c, -15) == d


def test_handle_api_matches_oneshot(oracle):
    from sharpziplib_b200 import datagen
    d = datagen.gen_xml(50000, 3).tobytes()
    df = oracle.Deflater(6, True)
    df.set_input(d)
    df.finish()
    out = b""
    while not df.finished:
        out += df.deflate(512)
    assert out == oracle.deflate(d, 6)
    assert df.total_in == len(d) and df.total_out == len(out)
    inf = oracle.Inflater(True)
    inf.set_input(out + b"TRAILER!")
    back = b""
    while not inf.finished:
        got = inf.inflate(4096)
        if not got:
            break
        back += got
    assert back == d and inf.remaining_input == 8 and inf.total_in == len(out)  # trap T14
