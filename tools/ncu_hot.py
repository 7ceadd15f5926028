#!/usr/bin/env python3
"""Top SASS lines by warp-stall samples from `ncu --page source --csv` output (development aid)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
c_src, c_s = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[hi + 1:]:
    try:
        data.append((float(r[c_s]), r))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1.0
top = sorted(data, key=lambda x: -x[0])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for s, r in top:
    st = sorted(((float(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{100 * s / tot:5.1f}%  {r[c_src][:90]:90s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")
agg = {}
for s, r in data:
    for i in stall_cols:
        agg[hdr[i][6:]] = agg.get(hdr[i][6:], 0.0) + float(r[i] or 0)
print("stall totals:", sorted(((round(v), k) for k, v in agg.items()), reverse=True)[:6])
