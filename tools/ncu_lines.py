"""Aggregates an ncu source page (cuda,sass) by CUDA-C line: share of samples, instructions executed, stall mix."""
import csv
import subprocess
import sys
from collections import defaultdict

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = next(i for i, r in enumerate(rows) if 'Address' in r and '# Samples' in r)
hdr = rows[h]
samp, ie, at = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
stall_cols = [i for i, c in enumerate(hdr) if c.startswith('stall_') and 'Not Issued' not in c]
agg = defaultdict(lambda: [0.0, 0, 0])
src = {}
st = defaultdict(float)
for r in rows[h + 1:]:
    if len(r) <= at or not r[0].strip().isdigit():
        continue
    l = int(r[0])
    src[l] = r[1]
    try:
        agg[l][0] += float(r[samp] or 0)
        agg[l][1] += int(r[ie] or 0)
        agg[l][2] += int(r[at] or 0)
    except ValueError:
        pass
    for c in stall_cols:
        try:
            st[hdr[c]] += float(r[c] or 0)
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values()) or 1
ti = sum(a[1] for a in agg.values())
print("kernel %s: %d samples, %d warp instructions, avg active threads %.1f" % (kernel, tot, ti, sum(a[2] for a in agg.values()) / max(1, ti)))
for l, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%5.1f%% L%-4d ie=%-9d thr=%4.1f  %s" % (100 * a[0] / tot, l, a[1], a[2] / max(1, a[1]), src[l][:105]))
print("stalls:", ", ".join("%s %.0f%%" % (k.replace('stall_', ''), 100 * v / tot) for k, v in sorted(st.items(), key=lambda x: -x[1])[:7]))
