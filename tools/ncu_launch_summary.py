"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/ncu_launch_summary.py launches.csv [title]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0]
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:48s} n={v[0]:4d} {v[1] / 1e3:9.3f} ms {100 * v[1] / tot:5.1f}%")
print(f"total {tot / 1e3:.3f} ms")
