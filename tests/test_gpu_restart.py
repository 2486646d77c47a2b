"""GPU tier: decoding a stream in pieces -- restart points of inflate plans and the Inflater handle fed at any
granularity (the reference's Inflater is a resumable mode machine, Inflater.cs:73-86 / :429-552; InflaterInputStream
hands it 4096 bytes at a time, Streams/InflaterInputStream.cs:486-498)."""
import io
import zlib

import numpy as np
import pytest

from sharpziplib_b200 import datagen

pytestmark = pytest.mark.gpu


def _streams(oracle):
    """(name, original, raw deflate stream) with several blocks of every type"""
    text = datagen.text_buffer(3, 600000).tobytes()
    mix = b"".join(datagen.silesia_mix(c, 90000, config=7).tobytes() for c in range(8))
    noise = datagen.Rng(99).bytes(150000).tobytes()
    out = [
        ("text_l6", text, oracle.deflate(text, level=6)),
        ("mix_l9", mix, oracle.deflate(mix, level=9)),
        ("mix_l1", mix, oracle.deflate(mix, level=1)),
        ("noise_l0_stored", noise, oracle.deflate(noise, level=0)),
        ("text_huffman_only", text[:200000], oracle.deflate(text[:200000], level=6, strategy=2)),
        ("text_flush_pattern", text[:300000], oracle.deflate(text[:300000], level=6, pattern=1)),
    ]
    co = zlib.compressobj(9, zlib.DEFLATED, -15)  # a foreign compressor: long blocks, static blocks for short tails
    out.append(("zlib_l9", mix, co.compress(mix) + co.flush()))
    return out


def _run_plan(z, comp_slices, out_caps, windows=None, start_bits=None):
    import torch
    n = len(comp_slices)
    dl = None if windows is None else [len(w) for w in windows]
    ip = z.InflatePlan([len(c) for c in comp_slices], out_caps, dict_lens=dl)
    if start_bits is not None:
        ip.set_start_bits(start_bits)
    hin = np.zeros(ip.in_bytes, dtype=np.uint8)
    for i in range(n):
        w = b"" if windows is None else windows[i]
        blob = w + comp_slices[i]
        o = int(ip.in_offsets[i])
        hin[o:o + len(blob)] = np.frombuffer(blob, dtype=np.uint8)
        assert int(ip.data_offsets[i]) == o + len(w)
    di = torch.from_numpy(hin).cuda()
    do = torch.zeros(ip.out_bytes, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(n, dtype=torch.int64, device="cuda")
    d_st = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_used = torch.zeros(n, dtype=torch.int64, device="cuda")
    ip.run(di, do, d_len, d_st, None, d_used)
    torch.cuda.synchronize()
    bit, pos = ip.restart_points()
    o2 = do.cpu().numpy()
    lens = d_len.cpu().numpy()
    outs = [o2[int(o):int(o) + int(l)].tobytes() for o, l in zip(ip.out_offsets, lens)]
    return outs, d_st.cpu().numpy(), d_used.cpu().numpy(), bit, pos


def test_restart_points_of_truncated_streams(z, oracle):
    """A stream cut anywhere reports the last block header it reached; a second plan that starts at that bit with the
    32 KiB of output in front of it as the window image produces exactly the rest."""
    rng = np.random.default_rng(7)
    cases = []
    for name, orig, comp in _streams(oracle):
        for cut in sorted(set([1, 2, len(comp) // 3, len(comp) - 1] + [int(x) for x in rng.integers(1, len(comp), 4)])):
            cases.append((name, orig, comp, cut))
    outs, st, used, bit, pos = _run_plan(z, [c[:cut] for _, _, c, cut in cases], [len(o) + 64 for _, o, _, _ in cases])
    second = []
    for i, (name, orig, comp, cut) in enumerate(cases):
        assert int(st[i]) & 0xFF == 8, (name, cut, st[i])  # B200Z_E_NEED_INPUT
        assert orig.startswith(outs[i]), (name, cut)
        assert 0 <= pos[i] <= len(outs[i]) and 0 <= bit[i] <= 8 * cut, (name, cut, bit[i], pos[i])
        second.append((int(bit[i]), int(pos[i])))
    # at least some of the cuts lie behind the first block
    assert sum(1 for b, p in second if p > 0) >= len(cases) // 3
    windows = [orig[max(0, p - 32768):p] for (name, orig, comp, cut), (b, p) in zip(cases, second)]
    slices = [comp[b >> 3:] for (name, orig, comp, cut), (b, p) in zip(cases, second)]
    outs2, st2, used2, bit2, pos2 = _run_plan(z, slices, [len(o) + 64 for _, o, _, _ in cases], windows=windows,
                                              start_bits=[b & 7 for b, p in second])
    for i, (name, orig, comp, cut) in enumerate(cases):
        b, p = second[i]
        assert int(st2[i]) == 0, (name, cut, st2[i])
        assert outs2[i] == orig[p:], (name, cut, b, p)
        assert (b >> 3) + int(used2[i]) == len(comp), (name, cut)


def test_restart_points_of_complete_streams(z, oracle):
    """For a complete stream the restart point is the header of the final block."""
    ss = _streams(oracle)
    outs, st, used, bit, pos = _run_plan(z, [c for _, _, c in ss], [len(o) + 64 for _, o, _ in ss])
    for i, (name, orig, comp) in enumerate(ss):
        assert int(st[i]) == 0 and outs[i] == orig
        # decoding on from there yields the tail and ends the stream
        b, p = int(bit[i]), int(pos[i])
        o2, st2, _, _, _ = _run_plan(z, [comp[b >> 3:]], [len(orig) + 64], windows=[orig[max(0, p - 32768):p]],
                                     start_bits=[b & 7])
        assert int(st2[0]) == 0 and o2[0] == orig[p:], name


def _feed(z, comp, chunk, nowrap=True, dictionary=None, out_chunk=1 << 16):
    inf = z.Inflater(nowrap)
    buf = bytearray(out_chunk)
    got = bytearray()
    pos = 0
    guard = 0
    while not inf.IsFinished:
        guard += 1
        assert guard < 200000, "no progress"
        if inf.IsNeedingDictionary:
            inf.SetDictionary(dictionary)
        n = inf.Inflate(buf)
        got += buf[:n]
        if n == 0 and inf.IsNeedingInput and not inf.IsFinished and not inf.IsNeedingDictionary:
            if pos >= len(comp):
                raise EOFError("Unexpected EOF")  # what InflaterInputStream.Fill raises (:494)
            inf.SetInput(comp[pos:pos + chunk])
            pos += chunk
    return bytes(got), inf, pos


@pytest.mark.parametrize("chunk", [517, 4096, 65536])
def test_inflater_handle_any_feed_granularity(z, oracle, chunk):
    for name, orig, comp in _streams(oracle):
        got, inf, pos = _feed(z, comp + b"TRAILING", chunk)
        assert got == orig, (name, chunk)
        fed = min(pos, len(comp) + 8)
        assert inf.TotalOut == len(orig) and inf.TotalIn == len(comp), (name, chunk)
        assert inf.RemainingInput == fed - len(comp), (name, chunk)  # exact: gzip / zip read their trailers from it


def test_inflater_handle_byte_by_byte(z, oracle):
    for n, level in ((0, 6), (1, 6), (300, 6), (1200, 1), (800, 0)):
        d = datagen.text_buffer(5, n).tobytes() if n else b""
        for nowrap in (True, False):
            c = oracle.deflate(d, level=level, nowrap=nowrap)
            got, inf, _ = _feed(z, c, 1, nowrap=nowrap)
            assert got == d and inf.IsFinished and inf.RemainingInput == 0 and inf.TotalIn == len(c), (n, level, nowrap)
            if not nowrap:
                assert inf.Adler == zlib.adler32(d)


def test_inflater_handle_zlib_and_dictionary_in_pieces(z, oracle):
    text = datagen.text_buffer(8, 400000).tobytes()
    dictionary = datagen.text_buffer(8, 50000).tobytes()  # longer than the window: only its tail matters
    co = zlib.compressobj(6, zlib.DEFLATED, 15, 9, zlib.Z_DEFAULT_STRATEGY, dictionary)
    c = co.compress(text) + co.flush()
    for chunk in (1000, 4096):
        got, inf, _ = _feed(z, c, chunk, nowrap=False, dictionary=dictionary)
        assert got == text and inf.Adler == zlib.adler32(text) and inf.RemainingInput == 0
    c2 = oracle.deflate(text, level=6, nowrap=False)
    got, inf, _ = _feed(z, c2, 777, nowrap=False, out_chunk=1000)  # small output buffers as well
    assert got == text and inf.Adler == zlib.adler32(text)
    bad = bytearray(c2)
    bad[-2] ^= 0x55
    with pytest.raises(z.SharpZipBaseException):
        _feed(z, bytes(bad), 5000, nowrap=False)  # "Adler chksum doesn't match"
    bad = bytearray(c2)
    bad[len(bad) // 2] ^= 0x10  # corrupt data in a later block: a decode error, the checksum, or a stream that never ends
    with pytest.raises((z.SharpZipBaseException, EOFError)):
        _feed(z, bytes(bad), 5000, nowrap=False)


def test_inflater_input_stream_default_buffer(z, oracle):
    """InflaterInputStream with its default 4096-byte buffer over a multi-megabyte stream: one pass over the data."""
    d = datagen.text_buffer(21, 3 << 20).tobytes()
    c = oracle.deflate(d, level=6, nowrap=False)
    ins = z.InflaterInputStream(io.BytesIO(c), z.Inflater(False))
    assert ins.read() == d
