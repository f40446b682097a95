#!/usr/bin/env python
"""Top source lines of an .ncu-rep by instructions executed (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; agg = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name") or r[0] == "": continue
    try: agg.append((cur, int(r[0]), r[1].strip(), int(r[7]), int(r[4])))
    except ValueError: pass
tot = sum(a[3] for a in agg); print("total warp-instructions", tot)
agg.sort(key=lambda a: -a[3])
for a in agg[:top]: print(f"{a[0]:18s} {a[1]:5d} {a[3]:12d} {a[3]/tot:6.3f} samp={a[4]:6d}  {a[2][:95]}")
