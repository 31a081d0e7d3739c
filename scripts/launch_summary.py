#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel totals and shares.  Usage: launch_summary.py in.csv out.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
acc = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("unnamed>::", "").strip()
    try:
        v = float(r[-1])
    except ValueError:
        continue
    v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(r[-2], 1e-6)
    a = acc.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in acc.values())
with open(sys.argv[2], "w") as f:
    f.write("kernel,launches,total_ms,share_pct\n")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{v[0]},{v[1]:.4f},{100 * v[1] / tot:.2f}\n")
    f.write(f"TOTAL,{sum(v[0] for v in acc.values())},{tot:.4f},100\n")
