#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import collections
import csv
import sys

src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui, mi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Metric Name")
agg = collections.OrderedDict()
n = 0
for r in data:
    if not r[mi].startswith("gpu__time_duration"):
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "us" else (v / 1e6 if r[ui] == "ns" else (v * 1e3 if r[ui] == "s" else v))
    c, t = agg.get(r[ki][:90], (0, 0.0))
    agg[r[ki][:90]] = (c + 1, t + v)
    n += 1
tot = sum(t for _, t in agg.values())
lines = [f"# {title}", f"# {n} launches captured, total {tot:.3f} ms (cold-cache, serialised: compare shares)", "kernel,launches,total_ms,share"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"\"{k}\",{c},{t:.3f},{t / tot:.4f}")
open(dst, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:12]))
