"""GPU tier, SURVEY.md row f4: the entry ciphers (csrc/b200z_crypto.cu) through the C-ABI against the oracle
(oracle/szl_crypto.cpp), the reference's own AES fixture and Python's zipfile."""
import io
import zipfile
import zlib

import numpy as np
import pytest

from test_oracle import AES_FIXTURE_TEXT, aes_fixture_entry, classic_zip

pytestmark = pytest.mark.gpu


def test_aes_reference_fixture_through_the_gpu(z):
    """test/.../Zip/ZipEncryptionHandling.cs:461-482 (ZipFileAESReadWithEmptyPassword): password verifier, auth code,
    AES-256-CTR and the inflate of the entry, all on the device"""
    from sharpziplib_b200.encryption import ZipAESTransform
    salt, pv, ct, mac, kb, method = aes_fixture_entry()
    t = ZipAESTransform("", salt, kb, False)
    assert t.PwdVerifier == pv
    plain = bytearray(len(ct))
    assert t.TransformBlock(ct, 0, len(ct), plain, 0) == len(ct)
    assert t.GetAuthCode()[:10] == mac and t.GetAuthCode()[:10] == mac
    out, used, status = z.inflate_batch([bytes(plain)], [len(AES_FIXTURE_TEXT) + 16])
    assert out[0] == AES_FIXTURE_TEXT
    t.Dispose()


def test_aes_7zip_fixture_through_the_gpu(z):
    """test/.../Zip/ZipFileHandling.cs:1707-1745 (ShouldReadAESBZip2ZipCreatedBy7Zip): password "password", AES-256; the payload is
    bzip2 (not this library's business: Python's bz2 reads it)"""
    import bz2
    from sharpziplib_b200.encryption import ZipAESTransform
    from test_oracle import aes_7zip_fixture_entry
    salt, pv, ct, mac, pw, text = aes_7zip_fixture_entry()
    t = ZipAESTransform(pw.decode(), salt, 32, False)
    plain = bytearray(len(ct))
    t.TransformBlock(ct, 0, len(ct), plain, 0)
    assert t.PwdVerifier == pv and t.GetAuthCode()[:10] == mac and bz2.decompress(bytes(plain)) == text
    t.Dispose()


@pytest.mark.parametrize("key_bytes", [16, 32])
def test_aes_transform_handle_matches_the_oracle(z, oracle, key_bytes):
    """TransformBlock in uneven pieces (the CTR position and the HMAC go on across calls), both directions"""
    from sharpziplib_b200.encryption import ZipAESTransform
    rng = np.random.default_rng(key_bytes)
    data = rng.bytes(70001)
    salt, pw = rng.bytes(key_bytes // 2), "pässword"
    want_ct, want_pv, want_mac = oracle.zip_aes(pw.encode(), salt, key_bytes, True, data)
    for write_mode, src, want in ((True, data, want_ct), (False, want_ct, data)):
        t = ZipAESTransform(pw, salt, key_bytes, write_mode)
        out = bytearray(len(src))
        pos = 0
        for piece in (1, 15, 16, 17, 47, 64, 500, 4096, 4097, 30000, 10 ** 9):
            k = min(piece, len(src) - pos)
            t.TransformBlock(src, pos, k, out, pos)
            pos += k
        assert pos == len(src) and bytes(out) == want
        assert t.PwdVerifier == want_pv and t.GetAuthCode() == want_mac
        t.Dispose()
    with pytest.raises(ValueError):
        ZipAESTransform("x", b"123", 24, True)  # ZipAESTransform.cs:43
    with pytest.raises(ValueError):
        ZipAESTransform("x", b"123", 16, True)  # :45


def test_aes_batch_many_entries(z, oracle):
    """derive_keys + one-shot batch: many entries of ragged sizes (0 .. 300 KB) at once, both key sizes"""
    from sharpziplib_b200 import encryption as E
    rng = np.random.default_rng(3)
    sizes = [0, 1, 15, 16, 17, 63, 64, 65, 119, 120, 1000, 4095, 4096, 65537, 300000] + [int(x) for x in rng.integers(1, 20000, 50)]
    for kb in (16, 32):
        bufs = [rng.bytes(n) for n in sizes]
        pws = [b"pw%d" % i + b"x" * (i % 90) for i in range(len(sizes))]  # some longer than SHA-1's block
        salts = [rng.bytes(kb // 2) for _ in sizes]
        keys = E.aes_derive_keys(pws, salts, kb)
        outs, auth = E.aes_batch(bufs, keys, kb, True)
        for i, b in enumerate(bufs):
            ct, pv, mac = oracle.zip_aes(pws[i], salts[i], kb, True, b)
            assert outs[i] == ct and keys[i, 2 * kb:].tobytes() == pv and auth[i].tobytes() == mac, (kb, i, sizes[i])
        back, auth2 = E.aes_batch(outs, keys, kb, False)
        assert back == bufs and (auth2 == auth).all()


def test_deflate_then_encrypt_matches_the_reference_order(z, oracle):
    """Streams/DeflaterOutputStream.cs:227-231: the codec's output goes through the transform -- here as two batch calls"""
    from sharpziplib_b200 import datagen, encryption as E
    bufs = [datagen.silesia_mix(i, 50000 + 999 * i, config=8).tobytes() for i in range(6)]
    comp, _ = z.deflate_batch(bufs, level=6)
    salts = [bytes([i] * 16) for i in range(6)]
    keys = E.aes_derive_keys([b"password"] * 6, salts, 32)
    enc, auth = E.aes_batch(comp, keys, 32, True)
    for i in range(6):
        ct, _, mac = oracle.zip_aes(b"password", salts[i], 32, True, oracle.deflate(bufs[i], level=6))
        assert enc[i] == ct and auth[i].tobytes() == mac


def test_pkzip_classic_matches_the_oracle_and_zipfile(z, oracle):
    from sharpziplib_b200 import encryption as E
    rng = np.random.default_rng(5)
    keys = E.PkzipClassic.GenerateKeys(b"secret")
    assert keys == oracle.pkzip_generate_keys(b"secret")
    with pytest.raises(ValueError):
        E.PkzipClassic.GenerateKeys(b"")  # PkzipClassic.cs:26-29
    data = rng.bytes(100001)
    crc = zlib.crc32(data)
    header = rng.bytes(11) + bytes([crc >> 24])
    enc_t = E.PkzipClassicEncryptCryptoTransform(keys)
    out = bytearray(12 + len(data))
    enc_t.TransformBlock(header, 0, 12, out, 0)  # the header first, then the data in pieces: the keys go on across calls
    pos = 0
    for piece in (1, 7, 4096, 10 ** 9):
        k = min(piece, len(data) - pos)
        enc_t.TransformBlock(data, pos, k, out, 12 + pos)
        pos += k
    want, _ = oracle.pkzip_transform(keys, True, header + data)
    assert bytes(out) == want
    assert zipfile.ZipFile(io.BytesIO(classic_zip(b"f.bin", bytes(out), crc, None))).read("f.bin", pwd=b"secret") == data
    dec_t = E.PkzipClassicDecryptCryptoTransform(keys)
    assert dec_t.TransformFinalBlock(bytes(out), 0, len(out)) == header + data
    with pytest.raises(z.InvalidOperationException):
        E.PkzipClassicDecryptCryptoTransform(b"short")  # PkzipClassic.cs:91-94
    # many entries at once
    bufs = [rng.bytes(int(n)) for n in rng.integers(0, 5000, 200)]
    ks = np.stack([np.frombuffer(oracle.pkzip_generate_keys(b"k%d" % i), np.uint8) for i in range(200)])
    outs, after = E.pkzip_batch(bufs, ks, True)
    for i, b in enumerate(bufs):
        w, ka = oracle.pkzip_transform(ks[i].tobytes(), True, b)
        assert outs[i] == w and after[i].tobytes() == ka
    back, _ = E.pkzip_batch(outs, ks, False)
    assert back == bufs


def test_deflate_reproduces_the_streams_the_reference_holds(z):
    """the GPU Deflater against the only emitted Deflater bytes in the reference tree (tests/test_oracle.py:
    reference_held_deflate_streams): SharpZipLib's own dynamic block for 56 bytes of text, and a static one"""
    from test_oracle import reference_held_deflate_streams
    (text, dyn), (small, stat) = reference_held_deflate_streams()
    for level in (1, 4, 5, 6, 9):
        outs, _ = z.deflate_batch([text, small], level=level)
        assert outs == [dyn, stat], level
    d = z.Deflater(6, True)
    d.SetInput(text)
    d.Finish()
    buf = bytearray(512)
    assert bytes(buf[:d.Deflate(buf)]) == dyn
