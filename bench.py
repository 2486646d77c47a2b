#!/usr/bin/env python
"""bench.py -- DEFLATE level-6 compress + inflate throughput (GB/s of uncompressed bytes) on B200.

One "step" = one pass of the hot path over one batch of synthetic input:
    deflate leg : config C3 -- raw Deflater level 6 on 1024 x 256 KiB Silesia-mix buffers
    inflate leg : config C2 -- raw Inflater on 256 x 1 MiB text buffers pre-deflated (level 6) by the oracle
value = (uncompressed bytes of both legs) / (device time of both legs), inputs resident in HBM.
e2e   = the same step through the public plan API from PINNED HOST buffers: H2D of the inputs, kernels, D2H of the
        produced sizes and bytes, all inside the timed region.
With --gpus N > 1 (torchrun, one rank per GPU) every rank runs the same shape on its own buffers (weak scaling), the
static Huffman tables are broadcast once over NCCL, and rank 0 reports total units / max-over-ranks time.
--impl reference times the CPU restatement of the reference (oracle/, all host threads) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "DEFLATE level-6 compress + inflate GB/s (uncompressed) at 1/2/4/8 B200 vs C# ref"
N_DEFLATE, SZ_DEFLATE = 1024, 262144      # C3
N_INFLATE, SZ_INFLATE = 256, 1 << 20       # C2


def _gen_deflate(i):
    from sharpziplib_b200 import datagen
    return datagen.silesia_mix(i, SZ_DEFLATE, config=3)


def _gen_inflate(i):
    from sharpziplib_b200 import datagen
    return datagen.text_buffer(i, SZ_INFLATE, config=2)


def make_inputs(rank, n_def, n_inf, workers):
    """(list of deflate inputs, list of inflate originals) as numpy uint8 arrays; buffer indices are offset per rank"""
    from concurrent.futures import ProcessPoolExecutor
    di = [rank * N_DEFLATE + i for i in range(n_def)]
    ii = [rank * N_INFLATE + i for i in range(n_inf)]
    if workers > 1:
        with ProcessPoolExecutor(max_workers=workers) as ex:
            d = list(ex.map(_gen_deflate, di, chunksize=8))
            t = list(ex.map(_gen_inflate, ii, chunksize=2))
    else:
        d = [_gen_deflate(i) for i in di]
        t = [_gen_inflate(i) for i in ii]
    return d, t


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """threads the reference arm may really use: logical CPUs, capped by the affinity mask and the cgroup CPU quota"""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
        except Exception:
            pass
    return n


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def captured_traffic(kernel, small):
    """dram bytes (read + write) of one launch of `kernel` on this workload, from the committed ncu --set full capture"""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if small or not os.path.exists(p):
        return None, None
    j = json.load(open(p))
    if j.get("kernel") != kernel:
        return None, None
    return int(j["dram_bytes_per_launch"]), j.get("source")


class CpuSample:
    """the oracle (C++ restatement of the reference) on a bounded sample of the same workload; only the C calls are timed"""

    def __init__(self, n_def, n_inf, d_inputs, comp_inf, inf_caps):
        import oracle_lib as O
        self.dj = O.BatchJob(0, [a.tobytes() for a in d_inputs[:n_def]], level=6)
        self.ij = O.BatchJob(1, comp_inf[:n_inf], out_caps=[c + 64 for c in inf_caps[:n_inf]])
        self.ud = int(self.dj.lens.sum())
        self.ui = int(sum(inf_caps[:n_inf]))

    def run(self, threads):
        t0 = time.perf_counter()
        self.dj.run(threads)
        t1 = time.perf_counter()
        self.ij.run(threads)
        t2 = time.perf_counter()
        return {"value": (self.ud + self.ui) / (t2 - t0) / 1e9, "deflate_gbs": self.ud / (t1 - t0) / 1e9,
                "inflate_gbs": self.ui / (t2 - t1) / 1e9, "seconds": t2 - t0}


def run_reference(args, rank, world, out):
    """--impl reference: the reference's CPU algorithm (oracle port; no .NET on the box) on all host threads"""
    if rank != 0:
        return
    import oracle_lib as O
    O.build()
    threads = host_threads()
    # bounded sample: 1/4 of each leg's buffers per step (64 MiB deflate + 64 MiB inflate of the same generators)
    n_def, n_inf = max(threads, N_DEFLATE // 4), max(min(threads, N_INFLATE), N_INFLATE // 4)
    n_def, n_inf = min(n_def, N_DEFLATE), min(n_inf, N_INFLATE)
    d, t = make_inputs(0, n_def, n_inf, min(threads, 32))
    comp = O.batch(0, [a.tobytes() for a in t], level=6, threads=threads)
    caps = [a.size for a in t]
    job = CpuSample(n_def, n_inf, d, comp, caps)
    for _ in range(args.warmup):
        job.run(threads)
    times, last = [], None
    for _ in range(args.steps):
        last = job.run(threads)
        times.append(last["seconds"])
    ub = n_def * SZ_DEFLATE + n_inf * SZ_INFLATE
    val = ub * len(times) / sum(times) / 1e9
    sample = "%d x 256 KiB deflate L6 + %d x 1 MiB inflate per step, %d threads" % (n_def, n_inf, threads)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C3 deflate L6 1024x256KiB + C2 inflate 256x1MiB (bounded CPU sample)", "sample": sample},
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample,
                             "note": "C++ restatement of SharpZipLib's managed path; no .NET runtime on the box"},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "deflate_gbs": last["deflate_gbs"], "inflate_gbs": last["inflate_gbs"], "gpu_launches": 0}
    out.write(json.dumps(line) + "\n")
    out.flush()


def _claim_stdout():
    """Libraries (NCCL's version banner, torchrun children) may print to fd 1; the contract is ONE JSON line on stdout.
    Everything else is redirected to stderr and the JSON line is written to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    out = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200z", choices=["b200z", "reference"])
    ap.add_argument("--small", action="store_true", help="1/8 size workload for quick checks (not a bench value)")
    ap.add_argument("--no-probe", dest="no_probe", action="store_true",
                    help="skip the informational probe of the opt-in search kernels that follows the measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world, out)
        return

    import torch
    import torch.distributed as dist
    import oracle_lib as O
    import sharpziplib_b200 as z
    from sharpziplib_b200.sharding import broadcast_static_tables
    torch.cuda.set_device(local_rank)
    z.init(local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        broadcast_static_tables(dist, device=torch.device("cuda", local_rank))  # the path's only collective
    n_def = N_DEFLATE // (8 if args.small else 1)
    n_inf = N_INFLATE // (8 if args.small else 1)
    ncpu = host_threads()
    workers = max(1, min(32, ncpu // max(1, world)))
    t_setup = time.time()
    d_np, t_np = make_inputs(rank, n_def, n_inf, workers)
    O.build()
    comp = O.batch(0, [a.tobytes() for a in t_np], level=6, threads=max(1, ncpu // max(1, world)))
    setup_s = time.time() - t_setup

    # ---- plans and resident device buffers ---------------------------------------------------------------
    dplan = z.DeflatePlan([a.size for a in d_np], level=6)
    iplan = z.InflatePlan([len(c) for c in comp], [a.size for a in t_np])
    h_din = torch.zeros(dplan.in_bytes, dtype=torch.uint8).pin_memory()
    for o, a in zip(dplan.in_offsets, d_np):
        h_din[o:o + a.size] = torch.from_numpy(a)
    h_iin = torch.zeros(iplan.in_bytes, dtype=torch.uint8).pin_memory()
    for o, c in zip(iplan.in_offsets, comp):
        h_iin[o:o + len(c)] = torch.frombuffer(bytearray(c), dtype=torch.uint8)
    dev = torch.device("cuda", local_rank)
    d_din = h_din.to(dev)
    d_iin = h_iin.to(dev)
    d_dout = torch.empty(dplan.out_bytes, dtype=torch.uint8, device=dev)
    d_iout = torch.empty(iplan.out_bytes, dtype=torch.uint8, device=dev)
    d_dlen = torch.zeros(n_def, dtype=torch.int64, device=dev)
    d_dst = torch.zeros(n_def, dtype=torch.int32, device=dev)
    d_ilen = torch.zeros(n_inf, dtype=torch.int64, device=dev)
    d_ist = torch.zeros(n_inf, dtype=torch.int32, device=dev)
    d_iused = torch.zeros(n_inf, dtype=torch.int64, device=dev)
    h_dout = torch.empty(dplan.out_bytes, dtype=torch.uint8).pin_memory()
    h_iout = torch.empty(iplan.out_bytes, dtype=torch.uint8).pin_memory()
    h_dlen = torch.zeros(n_def, dtype=torch.int64).pin_memory()
    h_ilen = torch.zeros(n_inf, dtype=torch.int64).pin_memory()
    U_def = sum(a.size for a in d_np)
    U_inf = sum(a.size for a in t_np)
    C_inf = sum(len(c) for c in comp)

    # The two legs are independent: the inflate leg (one long, latency-bound kernel on few warps) runs on a second stream
    # next to the deflate leg's kernels instead of after them.
    s_leg = torch.cuda.Stream(device=dev)
    ev_leg0, ev_leg1 = torch.cuda.Event(), torch.cuda.Event()

    def step_resident(overlap=True):
        cur = torch.cuda.current_stream()
        if not overlap:
            dplan.run(d_din, d_dout, d_dlen, d_dst)
            iplan.run(d_iin, d_iout, d_ilen, d_ist, None, d_iused)
            return
        # the inflate kernel starts when the match search (whose CTAs take a whole SM's shared memory) is through and
        # then shares the SMs with the parse / Huffman / bit-packing kernels
        dplan.run(d_din, d_dout, d_dlen, d_dst, stages=z.STAGE_SEARCH)
        ev_leg0.record(cur)
        s_leg.wait_event(ev_leg0)
        iplan.run(d_iin, d_iout, d_ilen, d_ist, None, d_iused, stream=s_leg)
        ev_leg1.record(s_leg)
        dplan.run(d_din, d_dout, d_dlen, d_dst, stages=z.STAGE_ENCODE)
        cur.wait_event(ev_leg1)

    # ---- end to end from pinned host buffers, public plan API.  Three streams (upload / kernels / download) and two sets
    # of device buffers, run as a software pipeline the way a server would: while step i's kernels run, step i+1's inputs
    # go up and step i-1's results come down.  Every step still uploads all of its inputs from pinned host memory and
    # reads all of its results (inflated bytes, packed deflate output, lengths) back inside the timed region; the last
    # step's downloads are drained before the closing timestamp.  The only host wait is for the packed size of the
    # PREVIOUS step (b200z_plan_pack: only the bytes actually produced cross PCIe).
    s_up, s_run, s_run2, s_down = (torch.cuda.Stream(device=dev) for _ in range(4))  # s_run2: the inflate leg

    class Slot:
        def __init__(self, first):
            self.d_iin = d_iin if first else torch.empty_like(d_iin)
            self.d_din = d_din if first else torch.empty_like(d_din)
            self.d_iout = d_iout if first else torch.empty_like(d_iout)
            self.d_dout = d_dout if first else torch.empty_like(d_dout)
            self.d_ilen, self.d_ist, self.d_iused = torch.zeros_like(d_ilen), torch.zeros_like(d_ist), torch.zeros_like(d_iused)
            self.d_dlen, self.d_dst = torch.zeros_like(d_dlen), torch.zeros_like(d_dst)
            self.d_pack = torch.empty(dplan.out_bytes, dtype=torch.uint8, device=dev)
            self.d_poff = torch.zeros(n_def + 1, dtype=torch.int64, device=dev)
            self.h_poff = torch.zeros(n_def + 1, dtype=torch.int64).pin_memory()
            self.h_dlen = torch.zeros(n_def, dtype=torch.int64).pin_memory()
            self.ev_hi, self.ev_hd, self.ev_i, self.ev_d, self.ev_dl, self.ev_s = (torch.cuda.Event() for _ in range(6))
            self.used = False

    slots = [Slot(True), Slot(False)]
    h_pack = torch.empty(dplan.out_bytes, dtype=torch.uint8).pin_memory()
    e2e_d2h = [0]
    pipe = {"k": 0, "pending": None, "last": None}

    def finish(sl):
        sl.ev_d.synchronize()  # the host needs the packed size before it can size the copy
        total = int(sl.h_poff[-1])
        with torch.cuda.stream(s_down):
            h_pack[:total].copy_(sl.d_pack[:total], non_blocking=True)
            sl.ev_dl.record(s_down)
        e2e_d2h[0] = total + h_iout.numel() + 8 * (2 * n_def + 1 + n_inf)
        pipe["last"] = sl

    def step_e2e():
        sl = slots[pipe["k"] & 1]
        pipe["k"] += 1
        if not sl.used:
            cur = torch.cuda.current_stream()
            for st in (s_up, s_run, s_run2, s_down):
                st.wait_stream(cur)
        with torch.cuda.stream(s_up):
            if sl.used:
                s_up.wait_event(sl.ev_d)   # the kernels that read this slot's inputs two steps ago are done
                s_up.wait_event(sl.ev_i)
            sl.d_iin.copy_(h_iin, non_blocking=True)
            sl.ev_hi.record(s_up)
            sl.d_din.copy_(h_din, non_blocking=True)
            sl.ev_hd.record(s_up)
        with torch.cuda.stream(s_run):
            if sl.used:
                s_run.wait_event(sl.ev_dl)  # this slot's previous outputs have left the device
            s_run.wait_event(sl.ev_hd)
            dplan.run(sl.d_din, sl.d_dout, sl.d_dlen, sl.d_dst, stream=s_run, stages=z.STAGE_SEARCH)
            sl.ev_s.record(s_run)
        with torch.cuda.stream(s_run2):
            if sl.used:
                s_run2.wait_event(sl.ev_dl)
            s_run2.wait_event(sl.ev_hi)  # free-running next to the deflate leg (gating it on ev_s measured slower here)
            iplan.run(sl.d_iin, sl.d_iout, sl.d_ilen, sl.d_ist, None, sl.d_iused, stream=s_run2)
            sl.ev_i.record(s_run2)
        with torch.cuda.stream(s_run):
            dplan.run(sl.d_din, sl.d_dout, sl.d_dlen, sl.d_dst, stream=s_run, stages=z.STAGE_ENCODE)
            dplan.pack(sl.d_dout, sl.d_dlen, sl.d_pack, sl.d_poff, stream=s_run)
            sl.h_poff.copy_(sl.d_poff, non_blocking=True)
            sl.h_dlen.copy_(sl.d_dlen, non_blocking=True)
            sl.ev_d.record(s_run)
        with torch.cuda.stream(s_down):
            s_down.wait_event(sl.ev_i)
            h_iout.copy_(sl.d_iout, non_blocking=True)
            h_ilen.copy_(sl.d_ilen, non_blocking=True)
        sl.used = True
        if pipe["pending"] is not None:
            finish(pipe["pending"])
        pipe["pending"] = sl

    def drain_e2e():
        if pipe["pending"] is not None:
            finish(pipe["pending"])
            pipe["pending"] = None
        cur = torch.cuda.current_stream()
        for st in (s_up, s_run, s_run2, s_down):
            cur.wait_stream(st)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, fin=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if fin is not None:
            fin()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    for _ in range(max(3, args.warmup)):
        step_resident()
    torch.cuda.synchronize()
    # ---- parity of what is being timed (not in the timed region) -------------------------------------------
    C_def = int(d_dlen.sum().item())
    assert int((d_dst != 0).sum().item()) == 0 and int((d_ist != 0).sum().item()) == 0, "device status != OK"
    lens = d_dlen.cpu().numpy()
    outb = d_dout.cpu().numpy()
    for i in range(0, n_def, max(1, n_def // 16)):
        ref = O.deflate(d_np[i].tobytes(), level=6)
        got = outb[dplan.out_offsets[i]:dplan.out_offsets[i] + lens[i]].tobytes()
        assert got == ref, "deflate parity failed for buffer %d" % i
    io = d_iout.cpu().numpy()
    for i in range(0, n_inf, max(1, n_inf // 16)):
        assert np.array_equal(io[iplan.out_offsets[i]:iplan.out_offsets[i] + t_np[i].size], t_np[i]), "inflate mismatch %d" % i

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    units = (U_def + U_inf) * world
    value = units / (ms_step / 1e3) / 1e9

    # ---- per-kernel device times for the roofline (separate steps, events between kernels) -------------------
    dplan.set_timing(True)
    iplan.set_timing(True)
    acc = {}
    reps = max(3, min(args.steps, 5))
    for _ in range(reps):
        step_resident(overlap=False)  # one kernel at a time: these are per-kernel durations
        torch.cuda.synchronize()
        for k, v in list(dplan.timings().items()) + list(iplan.timings().items()):
            acc[k] = acc.get(k, 0.0) + v / reps
    dplan.set_timing(False)
    iplan.set_timing(False)
    t_def = sum(v for k, v in acc.items() if k != "k_inflate")
    t_inf = acc.get("k_inflate", 0.0)
    dom = max(acc, key=acc.get)
    peak, peak_src = peaks()
    alg_bytes = (C_inf + U_inf) if dom == "k_inflate" else (U_def + C_def)
    achieved = alg_bytes / (acc[dom] / 1e3) / 1e9
    traffic, traffic_src = captured_traffic(dom, args.small)

    # ---- end to end from pinned host buffers ------------------------------------------------------------------
    for _ in range(3):
        step_e2e()
    drain_e2e()
    torch.cuda.synchronize()
    # what came back over PCIe is checked too
    h_poff, h_dlen = pipe["last"].h_poff, pipe["last"].h_dlen
    for i in range(0, n_def, max(1, n_def // 8)):
        got = h_pack[int(h_poff[i]):int(h_poff[i]) + int(h_dlen[i])].numpy().tobytes()
        assert got == O.deflate(d_np[i].tobytes(), level=6), "e2e deflate parity failed for buffer %d" % i
    hio = h_iout.numpy()
    for i in range(0, n_inf, max(1, n_inf // 8)):
        assert np.array_equal(hio[iplan.out_offsets[i]:iplan.out_offsets[i] + t_np[i].size], t_np[i]), "e2e inflate mismatch %d" % i
    n_e2e = max(3, args.steps)
    ms_e2e = timed(step_e2e, n_e2e, drain_e2e) / n_e2e
    e2e_val = units / (ms_e2e / 1e3) / 1e9

    # ---- CPU baseline: the oracle on one host core, bounded sample ----------------------------------------------
    cpu = None
    if rank == 0:
        ns_d, ns_i = n_def, n_inf  # the whole step once on one core: ~10 s (128 / 64 buffers took 1.4 s, too short a sample)
        r = CpuSample(ns_d, ns_i, d_np, comp, [a.size for a in t_np]).run(1)
        cpu = {"value": r["value"], "unit": "GB/s", "cores": 1, "kind": "port",
               "sample": "%d x 256 KiB deflate L6 + %d x 1 MiB inflate, single thread, %.1f s" % (ns_d, ns_i, r["seconds"]),
               "deflate_gbs": r["deflate_gbs"], "inflate_gbs": r["inflate_gbs"], "host_cores": ncpu,
               "note": "C++ restatement of SharpZipLib's managed path (oracle/); no .NET runtime on the box"}
    # ---- opt-in search kernels (csrc/experimental/k_tile_parse.cuh): probed AFTER every timed region, in a process of their
    # own (they had never run on a GPU when round 1 ended; a fault there must not touch this process).  Not part of
    # value / e2e / roofline: the line only carries what the probe printed, for the next round to start from.
    probe = None
    if rank == 0 and world == 1 and not args.small and not args.no_probe and not os.environ.get("B200Z_TILE_PARSE"):
        probe = {}
        t_probe = time.time()  # bounded: 90 s per variant, 150 s in all, and nothing after a variant that hung
        for variant in ("3", "4", "2", "1"):  # the simplest of the promising ones first: a hang ends the probe
            left = 150.0 - (time.time() - t_probe)
            if left < 20.0:
                probe["tile_parse" + variant] = {"ok": None, "skipped": "probe time budget spent"}
                continue
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "tile_parse_check.py"), variant, "256"], cwd=ROOT,
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=min(90.0, left))
                last = (r.stdout.decode(errors="replace").strip().splitlines() or [""])[-1]
                try:
                    probe["tile_parse" + variant] = json.loads(last)
                except ValueError:
                    probe["tile_parse" + variant] = {"ok": False, "rc": r.returncode, "stderr": r.stderr.decode(errors="replace")[-400:]}
            except subprocess.TimeoutExpired:
                probe["tile_parse" + variant] = {"ok": False, "timeout_s": round(min(90.0, left), 1)}
                t_probe = -1e9  # a hang: the other variants share most of the code, do not spend more of the run on them
            except Exception as e:  # the probe never costs the bench line
                probe["tile_parse" + variant] = {"ok": False, "error": repr(e)[:200]}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "C3 deflate L6 %dx256KiB + C2 inflate %dx1MiB per GPU" % (n_def, n_inf),
                       "l2": "inputs (256 MiB + 64 MiB compressed per leg) exceed the 126 MB L2; no flush needed",
                       "parity": "deflate bytes == oracle and inflate bytes == original, checked on 16 buffers each before timing",
                       "ratio_deflate": U_def / max(1, C_def), "ratio_inflate": U_inf / max(1, C_inf), "setup_s": setup_s,
                       # opt-in search kernel (csrc/experimental/k_tile_parse.cuh); "" = the default k_match + k_parse_chunk
                       "search_variant": ("tile_parse%s" % os.environ["B200Z_TILE_PARSE"]) if os.environ.get("B200Z_TILE_PARSE") in ("1", "2", "3", "4") else "default",
                       "link_run": int(os.environ.get("B200Z_LINK_RUN", "65536"))},
            "deflate_gbs": U_def / (t_def / 1e3) / 1e9 if t_def else None,
            "inflate_gbs": U_inf / (t_inf / 1e3) / 1e9 if t_inf else None,
            "kernels_ms": acc,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes": alg_bytes},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": "GB/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(dplan.in_bytes + iplan.in_bytes),
                    "d2h_bytes_per_step": int(e2e_d2h[0]),
                    "how": "pinned host buffers; upload / deflate-leg / inflate-leg / download streams over two device buffer sets (step i+1 uploads and step i-1 downloads overlap step i's kernels); packed D2H of the deflate output; %d steps + drain timed" % n_e2e},
            # kernels of libb200z.so launched inside the timed region of `value` (per step: k_links, k_match, 4 x k_parse_*,
            # k_plan, k_scan, k_emit, k_inflate)
            "gpu_launches": int(dplan.launches + iplan.launches) * args.steps,
            "gpu_launches_per_step": int(dplan.launches + iplan.launches),
            "clocks": clocks,
            # separate process, after the timed regions, 256 x 256 KiB: bit-exactness and per-kernel ms of the opt-in search
            # kernels next to the default path (tools/tile_parse_check.py); informational
            "experimental_probe": probe,
        }
        out.write(json.dumps(line) + "\n")
        out.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
