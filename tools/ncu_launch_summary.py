"""Summarises an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: launches, total / average ms, share."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Value" in r)
hdr = rows[h]
kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
        continue
    v = float(r[mv].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
    name = r[kn].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += v * scale
tot = sum(a[1] for a in agg.values())
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# serialised, cold-cache per-launch times; the kernel SHARES are what should agree with bench.py's live CUDA-event timing")
print("# %d launches, %.1f ms of kernel time" % (sum(a[0] for a in agg.values()), tot))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-60s launches=%4d total_ms=%10.3f avg_ms=%9.3f share=%5.1f%%" % (k[:60], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
