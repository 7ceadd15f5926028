#!/usr/bin/env python3
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (shares, not absolutes)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.OrderedDict(), 0.0
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    m = re.search(r"(k_[a-z0-9_]+|at::native::[a-zA-Z_]+|[a-zA-Z_]+_kernel[a-zA-Z_]*)", r[kn])
    name = m.group(1) if m else r[kn][:40]
    if r[mv].strip().lower() == "nan":  # kernels recorded while a CUDA graph is being captured do not run
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1000 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += v
    tot += v
print(f"{'kernel':34s} {'launches':>8s} {'total_us':>10s} {'share':>7s} {'avg_us':>8s}")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k[:34]:34s} {n:8d} {t:10.1f} {100 * t / tot:6.1f}% {t / n:8.1f}")
print(f"total_us {tot:.1f}  (serialised, cold-cache ncu timings: compare shares)")
