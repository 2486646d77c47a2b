"""GPU tier: levels 0-4 follow the reference for ANY call sequence -- SetInput schedules (SURVEY.md trap T9: DeflateStored
and DeflateFast depend on how the input arrives) and input after Flush() (their window-relative state is carried between
the segments of a stream).  The engine code is checked on the CPU by tests/test_cpu_model.py; here the kernels, the plans
and the Deflater handle run it."""
import io
import random
import zlib

import numpy as np
import pytest

from helpers import oracle_calls, random_chunks
from sharpziplib_b200 import datagen

pytestmark = pytest.mark.gpu


def _drain(d, size=65536):
    out = bytearray()
    buf = bytearray(size)
    while True:
        k = d.Deflate(buf)
        if k <= 0:
            break
        out += buf[:k]
    return bytes(out)


def _gpu_calls(z, level, segs, dictionary=None, busy_last=True):
    d = z.Deflater(level, dictionary is None)
    if dictionary is not None:
        d.SetDictionary(dictionary)
    out = bytearray()
    for i, (seg, chunks) in enumerate(segs):
        pos = 0
        for k, c in enumerate(chunks):
            d.SetInput(seg[pos:pos + c])
            pos += c
            if busy_last or k + 1 < len(chunks):
                out += _drain(d)
                assert d.IsNeedingInput
        d.Flush() if i + 1 < len(segs) else d.Finish()
        out += _drain(d, 4096)
    assert d.IsFinished
    return bytes(out)


@pytest.mark.parametrize("level", [0, 1, 2, 3, 4])
def test_deflater_any_call_sequence_levels_0_to_4(z, oracle, level):
    rnd = random.Random(2000 + level)
    text = datagen.gen_text(400000, 31).tobytes()
    mixed = datagen.silesia_mix(3, 300000, config=5).tobytes()
    for trial in range(8):
        src = text if trial % 2 == 0 else mixed
        n = rnd.choice([0, 5, 3000, 66000, 140000, 250000])
        nseg = rnd.choice([1, 1, 2, 3, 5])
        cuts = sorted(rnd.randrange(0, n + 1) for _ in range(nseg - 1))
        bounds = [0] + cuts + [n]
        segs = [(src[a:b], random_chunks(rnd, b - a)) for a, b in zip(bounds[:-1], bounds[1:])]
        # keep the number of device runs per trial bounded: the handle runs once per segment, chunks cost nothing
        busy_last = rnd.random() < 0.7
        dictionary = src[-rnd.choice([3, 500, 32506, 40000]):] if (level != 0 and rnd.random() < 0.3) else None
        ref = oracle_calls(level, segs, dictionary, busy_last)
        got = _gpu_calls(z, level, segs, dictionary, busy_last)
        assert got == ref, (level, trial, n, [len(s) for s, _ in segs], [c[:6] for _, c in segs], busy_last,
                            None if dictionary is None else len(dictionary))


@pytest.mark.parametrize("level", [0, 1, 4, 6])
def test_deflater_output_stream_small_writes(z, oracle, level):
    """DeflaterOutputStream.Write in 1000-byte pieces (each Write drains, Streams/DeflaterOutputStream.cs:506-510), then
    Flush + more writes + Finish: the bytes the reference's stream produces for the same calls."""
    data = datagen.silesia_mix(2, 200000, config=4).tobytes()
    ms = io.BytesIO()
    out = z.DeflaterOutputStream(ms, z.Deflater(level, True))
    out.IsStreamOwner = False
    for p in range(0, 120000, 1000):
        out.Write(data[p:p + 1000])
    out.Flush()
    for p in range(120000, len(data), 777):
        out.Write(data[p:p + 777])
    out.Finish()
    segs = [(data[:120000], [1000] * 120), (data[120000:], [len(data[p:p + 777]) for p in range(120000, len(data), 777)])]
    assert ms.getvalue() == oracle_calls(level, segs, None, True, nowrap=True)
    assert zlib.decompress(ms.getvalue(), -15) == data
    # one Write with everything: oracle's stream pattern
    ms2 = io.BytesIO()
    o2 = z.DeflaterOutputStream(ms2, z.Deflater(level, True))
    o2.IsStreamOwner = False
    for p in range(0, len(data), 4096):
        o2.Write(data[p:p + 4096])
    o2.Finish()
    assert ms2.getvalue() == oracle.deflate(data, level=level, pattern=2, chunk=4096)


@pytest.mark.parametrize("level", [0, 1, 3])
def test_deflate_plan_with_schedules(z, oracle, level):
    """batch plans: every stream with its own SetInput schedule (b200z_history.chunk_count / chunk_len)"""
    import torch
    rnd = random.Random(level)
    bufs = [datagen.silesia_mix(i, 30000 + 9000 * i, config=6).tobytes() for i in range(12)]
    steps = [rnd.choice([100, 1000, 4096, 32768, 65274, 70000]) for _ in bufs]
    chunk_lens = [[min(st, len(b) - p) for p in range(0, len(b), st)] for b, st in zip(bufs, steps)]
    plan = z.DeflatePlan([len(b) for b in bufs], level=level, chunk_lens=chunk_lens)
    host = np.zeros(plan.in_bytes, dtype=np.uint8)
    for o, b in zip(plan.in_offsets, bufs):
        host[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    d_in = torch.from_numpy(host).cuda()
    d_out = torch.zeros(plan.out_bytes, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(len(bufs), dtype=torch.int64, device="cuda")
    d_st = torch.zeros(len(bufs), dtype=torch.int32, device="cuda")
    plan.run(d_in, d_out, d_len, d_st)
    torch.cuda.synchronize()
    assert d_st.cpu().numpy().tolist() == [0] * len(bufs)
    lens = d_len.cpu().numpy()
    out = d_out.cpu().numpy()
    comp = [out[o:o + l].tobytes() for o, l in zip(plan.out_offsets, lens)]
    ref = [oracle.deflate(b, level=level, chunk=st) for b, st in zip(bufs, steps)]
    assert comp == ref
    # without a schedule the plan is "one SetInput with everything"
    outs, _ = z.deflate_batch(bufs, level=level)
    assert outs == [oracle.deflate(b, level=level) for b in bufs]


def test_engine_state_is_required_for_continued_fast_streams(z):
    """b200z_deflate_plan_create_ex: CONTINUE at levels 0-4 without the state of the previous segment is refused"""
    import ctypes as C
    from sharpziplib_b200 import _lib
    lens = np.array([1000], dtype=np.int64)
    hl = np.array([100], dtype=np.int64)
    pb = np.array([100], dtype=np.int64)
    for level in (0, 1, 4):
        hs = _lib.History(_lib.HIST_CONTINUE, 0, hl.ctypes.data, pb.ctypes.data, None, None)
        h = C.c_void_p()
        rc = _lib.lib().b200z_deflate_plan_create_ex(1, lens.ctypes.data, level, 0, 0, 0, C.addressof(hs), C.byref(h))
        assert rc == _lib.E_UNSUPPORTED
    assert _lib.lib().b200z_engine_state_bytes() == 2 * 65536 + 64
