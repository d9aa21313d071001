#!/usr/bin/env python
"""Aggregate an ncu SASS-page CSV by CUDA source line using nvdisasm -g line info.
usage: ncu_lines.py <sass_page.csv> <nvdisasm -g -c output> <kernel name substring> [ncars]"""
import csv, re, sys, collections
csvp, sassp, kern = sys.argv[1:4]
ncars = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
# address -> file:line from nvdisasm
amap, cur, inside = {}, None, False
for line in open(sassp):
    if line.startswith(".text.") and kern in line: inside = True; continue
    if inside and line.startswith("//---------------------"): break
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", line)
    if m: amap[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csvp)))
H = rows[1]
ai, ii, si = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples")
base = None
agg = collections.Counter(); smp = collections.Counter()
for r in rows[2:]:
    if len(r) != len(H): continue
    a = int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai])
    if base is None: base = a
    key = amap.get(a - base)
    agg[key] += float(r[ii]); smp[key] += float(r[si] or 0)
tot = sum(agg.values()); ts = sum(smp.values())
print(f"total warp-instructions {tot:.3e}  ({tot / ncars:.0f} per car), samples {ts:.0f}")
lines = {}
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:60]:
    f, ln = k if k else ("?", 0)
    try:
        src = open(f"/root/repo/ft_grandprix_b200/csrc/{f}").read().split("\n")[ln - 1].strip()[:90]
    except Exception:
        src = ""
    print(f"{v / tot * 100:5.1f}% inst {smp[k] / ts * 100:5.1f}% smp {v / ncars:7.0f}/car  {f}:{ln}  {src}")
