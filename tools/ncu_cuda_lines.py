#!/usr/bin/env python
"""Per-source-line totals from `ncu -i X --page source --print-source cuda,sass --csv`."""
import csv, sys, collections
rows = csv.reader(open(sys.argv[1]))
ncars = float(sys.argv[2]) if len(sys.argv) > 2 else 1
f = None; H = None; agg = collections.Counter(); smp = collections.Counter(); txt = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": f = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": H = r; ii = H.index("Instructions Executed"); si = H.index("# Samples"); continue
    if H and len(r) == len(H) and r[0]:
        try:
            agg[(f, int(r[0]))] += float(r[ii]); smp[(f, int(r[0]))] += float(r[si]); txt[(f, int(r[0]))] = r[1].strip()[:95]
        except ValueError:
            pass
tot = sum(agg.values()); ts = sum(smp.values()) or 1
print(f"total {tot:.3e} warp-instr ({tot / ncars:.0f}/car)")
for k, v in agg.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 40):
    print(f"{v / tot * 100:5.1f}% inst {smp[k] / ts * 100:5.1f}% smp {v / ncars:8.0f}/car {k[0]}:{k[1]} {txt[k]}")
