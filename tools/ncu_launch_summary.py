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
# kernels that run once per model load (weight folds) or once per generate() call (time vectors), not per denoising step
ONCE = ("matmul_f64_kernel", "small_linear_kernel", "clip_embed_kernel")
step = {k: v for k, v in agg.items() if not any(o in k for o in ONCE)}
once = {k: v for k, v in agg.items() if any(o in k for o in ONCE)}
tot = sum(v[1] for v in step.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
for k, v in sorted(step.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:48s} n={v[0]:4d} {v[1] / 1e3:9.3f} ms {100 * v[1] / tot:5.1f}%")
print(f"total {tot / 1e3:.3f} ms")
if once:
    print("once per model load / per generate() call (not part of a denoising step):")
    for k, v in sorted(once.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:48s} n={v[0]:4d} {v[1] / 1e3:9.3f} ms")
