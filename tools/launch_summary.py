#!/usr/bin/env python
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    agg.setdefault(r[ki].split("(")[0][:64], []).append(v)
tot = sum(sum(v) for v in agg.values())
print(f"{'kernel':64s} {'launches':>8s} {'avg us':>10s} {'total us':>11s} {'share':>6s}")
for k, v in agg.items():
    print(f"{k:64s} {len(v):8d} {sum(v)/len(v):10.1f} {sum(v):11.1f} {sum(v)/tot:6.3f}")
