"""Checks the opt-in search kernels (csrc/experimental/k_tile_parse.cuh, B200Z_TILE_PARSE=1|2|3|4) on cuda:0 against the oracle and
times them next to the default path (k_match + k_parse_chunk).  Run in a process of its own: the kernels had never run on a
GPU when this was written, and a faulting kernel takes the CUDA context of its process with it.

    python tools/tile_parse_check.py [variant=2] [n_buffers=64]      ->  one JSON line on stdout, exit 0 iff bit-exact

tests/test_gpu_tile_parse.py calls it in a subprocess (GPU tier); with a budget it is the first thing to run in round 2:
    B200Z_TILE_PARSE=2 python bench.py        # the bench line then carries config.search_variant = "tile_parse2"
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINK_RUN = "131072"  # B200Z_LINK_RUN of the variant run on the C3 shape (default 65536: 37 % warm-up steps on 256 KiB buffers, here 12 %)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def per_kernel_ms(z, bufs, level, reps=3):
    """device time per kernel of one deflate plan over bufs (events between the kernels), averaged"""
    import numpy as np
    import torch
    from sharpziplib_b200 import batch
    plan = batch.DeflatePlan([len(b) for b in bufs], level=level)
    d_in = torch.zeros(plan.in_bytes, dtype=torch.uint8, device="cuda")
    host = np.zeros(plan.in_bytes, dtype=np.uint8)
    for i, b in enumerate(bufs):
        o = int(plan.data_offsets[i])
        host[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
    d_in.copy_(torch.from_numpy(host))
    d_out = torch.zeros(plan.out_bytes, dtype=torch.uint8, device="cuda")
    d_len = torch.zeros(len(bufs), dtype=torch.int64, device="cuda")
    d_st = torch.zeros(len(bufs), dtype=torch.int32, device="cuda")
    plan.run(d_in, d_out, d_len, d_st)  # warm-up
    torch.cuda.synchronize()
    plan.set_timing(True)
    acc = {}
    for _ in range(reps):
        plan.run(d_in, d_out, d_len, d_st)
        torch.cuda.synchronize()
        for k, v in plan.timings().items():
            acc[k] = acc.get(k, 0.0) + v / reps
    plan.set_timing(False)
    lens = d_len.cpu().numpy()
    out = d_out.cpu().numpy()
    outs = [out[int(plan.out_offsets[i]):int(plan.out_offsets[i]) + int(lens[i])].tobytes() for i in range(len(bufs))]
    assert int((d_st != 0).sum().item()) == 0, "device status != OK"
    plan.close()
    return acc, outs


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else "2"
    nbuf = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    assert variant in ("1", "2", "3", "4")
    import sharpziplib_b200 as z
    from sharpziplib_b200 import datagen
    import oracle_lib as O
    from helpers import corpus_small, crafted_t8
    O.build()
    z.init(0)
    res = {"variant": "tile_parse" + variant, "ok": False}
    t0 = time.time()
    # ---- bit-exactness: small corpus (empty, tiny, sub-tile), multi-tile streams, window slides (T8), levels 5/6/9 --------
    os.environ["B200Z_TILE_PARSE"] = variant
    names, small = zip(*corpus_small())
    multi = [crafted_t8(), crafted_t8(40000), bytes(200000), b"abcdefgh" * 40000] + \
        [datagen.silesia_mix(c, 300000 + 12345 * c, config=7).tobytes() for c in (0, 3, 4, 6)]
    for level in (5, 6, 9):
        outs, _ = z.deflate_batch(list(small) + multi, level=level)
        for d, o in zip(list(small) + multi, outs):
            if o != O.deflate(d, level=level):
                res["mismatch"] = {"level": level, "len": len(d)}
                print(json.dumps(res))
                return 1
    outs, _ = z.deflate_batch(list(small), level=6, strategy=1)
    assert all(o == O.deflate(d, level=6, strategy=1) for d, o in zip(small, outs)), "Filtered strategy"
    outs, _ = z.deflate_batch(list(small), level=6, end_mode=1)
    assert all(o == O.deflate(d, level=6, pattern=1) for d, o in zip(small, outs)), "Flush -> Finish"
    # ---- the C3 shape: nbuf x 256 KiB of the Silesia mix, level 6; per-kernel device times of both paths ---------------
    bufs = [datagen.silesia_mix(i, 262144).tobytes() for i in range(nbuf)]
    refs = O.batch(0, bufs, level=6, threads=8)
    os.environ["B200Z_LINK_RUN"] = LINK_RUN  # the variant runs also time k_links with longer runs (its own kernel, its own interval)
    t_var, outs = per_kernel_ms(z, bufs, 6)
    del os.environ["B200Z_LINK_RUN"]
    if outs != refs:
        res["mismatch"] = {"level": 6, "shape": "c3"}
        print(json.dumps(res))
        return 1
    del os.environ["B200Z_TILE_PARSE"]
    t_def, outs = per_kernel_ms(z, bufs, 6)
    assert outs == refs, "default path"
    # ---- level 9 (chain 4096: k_match walks 24.6x the candidates the reference walks there), 64 buffers ------------------
    b9 = bufs[:64]
    refs9 = O.batch(0, b9, level=9, threads=8)
    os.environ["B200Z_TILE_PARSE"] = variant
    t9_var, outs = per_kernel_ms(z, b9, 9, reps=2)
    if outs != refs9:
        res["mismatch"] = {"level": 9, "shape": "c3"}
        print(json.dumps(res))
        return 1
    del os.environ["B200Z_TILE_PARSE"]
    t9_def, outs = per_kernel_ms(z, b9, 9, reps=2)
    assert outs == refs9, "default path, level 9"
    res["level9_search_plus_parse_ms"] = {"buffers": len(b9), "default": t9_def.get("k_match", 0.0) + t9_def.get("k_parse", 0.0),
                                          "variant": t9_var.get("k_tile_parse", 0.0) + t9_var.get("k_parse", 0.0)}
    search_def = t_def.get("k_match", 0.0) + t_def.get("k_parse", 0.0)
    search_var = t_var.get("k_tile_parse", 0.0) + t_var.get("k_parse", 0.0)
    res.update({"ok": True, "buffers": nbuf, "bytes": sum(len(b) for b in bufs), "default_ms": t_def, "variant_ms": t_var,
                "search_plus_parse_ms": {"default": search_def, "variant": search_var},
                "k_links_ms": {"run_65536": t_def.get("k_links"), "run_" + LINK_RUN: t_var.get("k_links")},
                "speedup_search_plus_parse": (search_def / search_var) if search_var else None, "seconds": time.time() - t0})
    print(json.dumps(res))
    return 0


if __name__ == "__main__":
    sys.exit(main())
