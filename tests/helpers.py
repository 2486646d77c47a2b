import numpy as np

from sharpziplib_b200 import datagen


def crafted_t8(n_extra=400):
    """Input that makes position 65273 (first slide position, trap T8) a loop top whose only chain candidate sits at
    distance exactly 32506: the window-index-0 sentinel must reject it."""
    rng = datagen.Rng(0x7E57)
    a = rng.bytes(65273 + n_extra)
    # make bytes incompressible noise, then plant an 8-byte marker at 65273-32506 and again at 65273
    marker = bytes([1, 2, 3, 250, 251, 252, 253, 254])
    p = 65273
    q = p - 32506
    a[q:q + 8] = np.frombuffer(marker, dtype=np.uint8)
    a[p:p + 8] = np.frombuffer(marker, dtype=np.uint8)
    return a.tobytes()


def corpus_small():
    """(name, bytes) pairs the oracle finishes in well under a second each"""
    out = []
    for n in (0, 1, 2, 3, 4, 10, 100, 1000, 4096, 20000, 70000):
        for cls in range(8):
            if n < 100 and cls not in (0, 7):
                continue
            out.append(("%s_%d" % (datagen.CLASS_NAMES[cls], n), datagen.silesia_mix(cls, n, config=9).tobytes() if n else b""))
    out.append(("zeros_100000", bytes(100000)))
    out.append(("ab_70000", b"ab" * 35000))
    out.append(("t8", crafted_t8()))
    return out


def oracle_calls(level, segs, dictionary=None, busy_last=True, nowrap=None):
    """reference call sequence over oracle.Deflater: per segment SetInput(chunk) + Deflate-until-needs-input for every
    chunk (the last one only if busy_last), then Flush() (Finish() for the last segment) and drain"""
    from oracle_lib import Deflater
    d = Deflater(level, nowrap=(dictionary is None) if nowrap is None else nowrap)
    if dictionary is not None:
        d.set_dictionary(dictionary)
    out = bytearray()

    def drain():
        while True:
            b = d.deflate(4096)
            if not b:
                break
            out.extend(b)
    for i, (seg, chunks) in enumerate(segs):
        pos = 0
        for k, c in enumerate(chunks):
            d.set_input(seg[pos:pos + c])
            pos += c
            if busy_last or k + 1 < len(chunks):
                drain()
                assert d.needs_input
        if i + 1 < len(segs):
            d.flush()
        else:
            d.finish()
        drain()
    return bytes(out)


def random_chunks(rnd, n):
    if n == 0:
        return [0] if rnd.random() < 0.5 else []
    style = rnd.choice(["one", "small", "mixed", "window"])
    if style == "one":
        return [n]
    out, left = [], n
    while left:
        if style == "small":
            c = rnd.choice([1, 2, 3, 100, 512, 4096])
        elif style == "mixed":
            c = rnd.randrange(1, 70000)
        else:
            c = rnd.choice([32767, 32768, 65273, 65274, 65275, 262, 261, 1])
        c = min(c, left)
        out.append(c)
        left -= c
    return out


