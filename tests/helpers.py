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
