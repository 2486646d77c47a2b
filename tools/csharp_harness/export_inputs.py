"""Writes the generator's buffers (configs C1, C3 sample, C5 sample) as <dir>/<name>.bin for the C# harness,
and -- the other direction -- checks a reference_digests.json produced by it against the oracle:
    python tools/csharp_harness/export_inputs.py export /tmp/szl_inputs
    python tools/csharp_harness/export_inputs.py check tests/golden/reference_digests.json"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from sharpziplib_b200 import datagen  # noqa: E402


def buffers():
    yield "c1_text_4k", datagen.text_buffer(0, 4096, config=1).tobytes()
    for i in range(16):
        yield "c3_mix_%02d_256k" % i, datagen.silesia_mix(i, 262144, config=3).tobytes()
    for i, size in enumerate((4096, 16384, 65536, 262144, 1 << 20, 4 << 20)):
        yield "c5_mix_%d" % size, datagen.silesia_mix(i, size, config=5).tobytes()
    yield "c2_text_1m", datagen.text_buffer(0, 1 << 20, config=2).tobytes()


if __name__ == "__main__":
    if sys.argv[1] == "export":
        os.makedirs(sys.argv[2], exist_ok=True)
        for name, b in buffers():
            open(os.path.join(sys.argv[2], name + ".bin"), "wb").write(b)
            print(name, len(b))
    else:
        import oracle_lib as O
        ref = json.load(open(sys.argv[2]))
        bad = 0
        for name, b in buffers():
            e = ref["buffers"].get(name + ".bin")
            if not e:
                continue
            assert hashlib.sha256(b).hexdigest() == e["in_sha256"], name + ": the generator's bytes differ"
            got = hashlib.sha256(O.deflate(b, level=ref["level"])).hexdigest()
            print(name, "OK" if got == e["out_sha256"] else "ORACLE != REFERENCE")
            bad += got != e["out_sha256"]
        sys.exit(1 if bad else 0)
