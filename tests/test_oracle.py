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


# ---- entry ciphers (oracle/szl_crypto.cpp) -------------------------------------------------------
# The AES-encrypted archive the reference's own tests hold (test/.../Zip/ZipEncryptionHandling.cs:452-456: one entry "test",
# AES-256, deflated, empty password; :461-482 expects the text below).
AES_FIXTURE_B64 = GOLD["aes_empty_password_zip"]["zip_base64"]  # extracted from the reference's test source by tests/golden/make_golden.py
AES_FIXTURE_TEXT = GOLD["aes_empty_password_zip"]["text"].encode()


def aes_fixture_entry():
    """(salt, password verifier, ciphertext, 10-byte auth code, key bytes, compression method) of the fixture's entry"""
    import base64
    import struct
    z = base64.b64decode("".join(AES_FIXTURE_B64.split()))
    sig, _, flags, method, _, _, _, _, _, nl, xl = struct.unpack("<IHHHHHIIIHH", z[:30])
    assert sig == 0x04034B50 and method == 99 and flags & 1
    extra, i, csize, strength, real = z[30 + nl:30 + nl + xl], 0, None, None, None
    while i < len(extra):
        tag, sz = struct.unpack("<HH", extra[i:i + 4])
        body = extra[i + 4:i + 4 + sz]
        i += 4 + sz
        if tag == 1:
            _, csize = struct.unpack("<QQ", body[:16])
        if tag == 0x9901:
            _, _, strength, real = struct.unpack("<H2sBH", body)
    kb = {1: 16, 3: 32}[strength]
    data = z[30 + nl + xl:30 + nl + xl + csize]
    return data[:kb // 2], data[kb // 2:kb // 2 + 2], data[kb // 2 + 2:-10], data[-10:], kb, real


def test_crypto_oracle_known_answers(oracle):
    O = oracle
    k16, k32, pt = bytes(range(16)), bytes(range(32)), bytes.fromhex("00112233445566778899aabbccddeeff")
    assert O.aes_encrypt_block(k16, pt).hex() == "69c4e0d86a7b0430d8cdb78070b4c55a"  # FIPS 197 C.1
    assert O.aes_encrypt_block(k32, pt).hex() == "8ea2b7ca516745bfeafc49904b496089"  # FIPS 197 C.3
    assert O.sha1(b"abc").hex() == "a9993e364706816aba3e25717850c26c9cd0d89d"  # RFC 3174
    assert O.sha1(b"").hex() == "da39a3ee5e6b4b0d3255bfef95601890afd80709"
    assert O.pbkdf2_sha1(b"password", b"salt", 4096, 20).hex() == "4b007901b765489abead49d926f721d065a429c1"  # RFC 6070
    assert O.hmac_sha1(b"Jefe", b"what do ya want for nothing?").hex() == "effcdf6ae5eb2fa2d27416d5f184df9c259a7c79"  # RFC 2202


def test_crypto_oracle_reads_the_reference_aes_fixture(oracle):
    O = oracle
    salt, pv, ct, mac, kb, method = aes_fixture_entry()
    plain, verifier, auth = O.zip_aes(b"", salt, kb, False, ct)
    assert verifier == pv and auth[:10] == mac and method == 8
    assert O.inflate(plain, nowrap=True)[0] == AES_FIXTURE_TEXT


def aes_7zip_fixture_entry():
    """the entry of the AES-256 + bzip2 archive written by 7-zip that the reference's tests hold (ZipFileHandling.cs:1707-1745)"""
    import base64
    import struct
    g = GOLD["aes_bzip2_zip_by_7zip"]
    z = base64.b64decode(g["zip_base64"])
    sig, _, flags, method, _, _, _, csize, _, nl, xl = struct.unpack("<IHHHHHIIIHH", z[:30])
    assert sig == 0x04034B50 and method == 99 and flags & 1
    data = z[30 + nl + xl:30 + nl + xl + csize]
    return data[:16], data[16:18], data[18:-10], data[-10:], g["password"].encode(), g["text"].encode()


def test_crypto_oracle_reads_the_reference_7zip_aes_fixture(oracle):
    import bz2
    salt, pv, ct, mac, pw, text = aes_7zip_fixture_entry()
    plain, verifier, auth = oracle.zip_aes(pw, salt, 32, False, ct)
    assert verifier == pv and auth[:10] == mac and bz2.decompress(plain) == text


def test_crypto_oracle_against_an_independent_library(oracle):
    O = oracle
    hz = pytest.importorskip("cryptography.hazmat.primitives")
    from cryptography.hazmat.primitives import hashes, hmac
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    from cryptography.hazmat.primitives.kdf.pbkdf2 import PBKDF2HMAC
    rng = np.random.default_rng(7)
    for bs in (16, 32):
        for n in (0, 1, 15, 16, 17, 63, 64, 65, 4096, 100003):
            pw, salt, data = b"pass\xc3\xa9word", rng.bytes(bs // 2), rng.bytes(n)
            kb = PBKDF2HMAC(hashes.SHA1(), 2 * bs + 2, salt, 1000).derive(pw)
            enc = Cipher(algorithms.AES(kb[:bs]), modes.ECB()).encryptor()
            ks = b"".join(enc.update((i + 1).to_bytes(16, "little")) for i in range((n + 15) // 16))
            ct = bytes(a ^ b for a, b in zip(data, ks))
            h = hmac.HMAC(kb[bs:2 * bs], hashes.SHA1())
            h.update(ct)
            mac = h.finalize()
            assert O.zip_aes(pw, salt, bs, True, data, piece=777) == (ct, kb[2 * bs:], mac)
            assert O.zip_aes(pw, salt, bs, False, ct) == (data, kb[2 * bs:], mac)


def classic_zip(name, payload, crc, keys_after_header_fn):
    """a one-entry stored archive with PKZIP classic encryption (flag bit 0), header + data already encrypted in `payload`"""
    import struct
    lh = struct.pack("<IHHHHHIIIHH", 0x04034B50, 20, 1, 0, 0, 0x21, crc, len(payload), len(payload) - 12, len(name), 0) + name
    cd = struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 20, 20, 1, 0, 0, 0x21, crc, len(payload), len(payload) - 12, len(name), 0, 0, 0, 0, 0,
                     0) + name
    eocd = struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, 1, 1, len(cd), len(lh) + len(payload), 0)
    return lh + payload + cd + eocd


def test_pkzip_classic_oracle_is_read_by_python_zipfile(oracle):
    O = oracle
    import io
    import zipfile
    import zlib
    data = bytes(range(256)) * 37 + b"tail"
    crc = zlib.crc32(data)
    keys = O.pkzip_generate_keys(b"secret")
    header = bytes(range(11)) + bytes([crc >> 24])  # ZipOutputStream.WriteEncryptionHeader: 11 random bytes + the CRC's top byte
    enc, keys_after = O.pkzip_transform(keys, True, header + data)
    zf = zipfile.ZipFile(io.BytesIO(classic_zip(b"f.bin", enc, crc, None)))
    assert zf.read("f.bin", pwd=b"secret") == data
    dec, keys_after2 = O.pkzip_transform(keys, False, enc)
    assert dec == header + data and keys_after2 == keys_after


def reference_held_deflate_streams():
    """(input, raw deflate stream) pairs the reference's own tests hold: the payload inside its AES-encrypted archive (decrypted
    with the oracle; written by SharpZipLib itself -- a DYNAMIC block for 56 bytes of text, where zlib emits a static one) and
    the entry of TestFileBadCDGoodCD64"""
    salt, pv, ct, mac, kb, method = aes_fixture_entry()
    import oracle_lib
    plain, verifier, auth = oracle_lib.zip_aes(b"", salt, kb, False, ct)
    assert verifier == pv and auth[:10] == mac and method == 8
    raw2 = bytes.fromhex(GOLD["inflate_ok"]["raw_hex"])
    return [(AES_FIXTURE_TEXT, plain), (oracle_lib.inflate(raw2, nowrap=True)[0], raw2)]


def test_deflater_bytes_against_streams_the_reference_holds(oracle):
    """The only emitted Deflater bytes the reference tree holds.  They pin the Huffman half of the oracle's Deflater (Tree.BuildTree /
    BuildLength / BuildCodes with the dummy codes of trap T3, SendAllTrees / WriteTree, the dynamic / static / stored decision of
    FlushBlock, CompressBlock, the bit writer) against real SharpZipLib output; neither input has a repeated trigram, so match
    finding stays unpinned."""
    (text, dyn), (small, stat) = reference_held_deflate_streams()
    assert dyn[0] & 7 == 0b101  # BFINAL = 1, BTYPE = 2: a dynamic block
    for level in range(1, 10):
        for strategy in (0, 1, 2):
            assert oracle.deflate(text, level=level, strategy=strategy) == dyn, (level, strategy)
        assert oracle.deflate(small, level=level) == stat
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        assert c.compress(text) + c.flush() != dyn  # zlib would not have written this stream: it is SharpZipLib's
