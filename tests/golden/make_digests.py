"""Pins the synthetic generator and the oracle's level-6 output for 8 x 64 KiB (one buffer per data class) as SHA-256
digests, so that a different box (the GPU box) can check it regenerates the same inputs and the same oracle bytes.
Self-consistency only: these are NOT reference-produced vectors (the reference cannot run here; DESIGN.md "Oracle")."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
from sharpziplib_b200 import datagen  # noqa: E402

out = {"datagen_sha256_first_64k": {}, "oracle_deflate_l6_sha256_64k": {}}
for i in range(8):
    d = datagen.silesia_mix(i, 65536).tobytes()
    out["datagen_sha256_first_64k"][datagen.CLASS_NAMES[i]] = hashlib.sha256(d).hexdigest()
    out["oracle_deflate_l6_sha256_64k"][datagen.CLASS_NAMES[i]] = hashlib.sha256(O.deflate(d, level=6)).hexdigest()
json.dump(out, open(os.path.join(HERE, "oracle_digests.json"), "w"), indent=1)
print("wrote oracle_digests.json")
