#!/usr/bin/env python
"""Attribute an ncu source-page CSV of the step kernel (ncu -i rep --page source --csv) to the functions of
mushr_step_quad.cuh using nvdisasm -g line info of the same cubin.
usage: ncu_regions.py <sass_page.csv> <nvdisasm -g -c output> [launch index]"""
import csv, re, collections, sys, os
csvp, disp = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
amap = {}; cur = None; inside = False
for line in open(disp):
    if line.startswith('.text.'):
        inside = 'step_quad_kernel' in line
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', line)
    if m: amap[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csvp)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
H = rows[hdr[which]]
end = hdr[which + 1] - 2 if which + 1 < len(hdr) else len(rows)
ai, si, ii, ni = H.index('Address'), H.index('# Samples'), H.index('Instructions Executed'), H.index('stall_no_inst')
stalls = [h for h in H if h.startswith('stall_') and 'Not Issued' not in h]
src = open(os.path.join(ROOT, 'ft_grandprix_b200', 'csrc', 'mushr_step_quad.cuh')).read().split('\n')
marks = []
for i, line in enumerate(src, 1):
    m = re.match(r'^(FT_QN|FT_HD|FT_HDN|FT_STEP)\s+\w+\s+(\w+)\(', line)
    if m: marks.append((m.group(2), i))
REG = [(n, l, (marks[k + 1][1] if k + 1 < len(marks) else 10 ** 9)) for k, (n, l) in enumerate(marks)]
def region(k):
    if k is None: return '?'
    f, l = k
    if f == 'mushr_step_quad.cuh':
        for name, lo, hi in REG:
            if lo - 3 <= l < hi - 3: return name
    return f
inst = collections.Counter(); smp = collections.Counter(); st = collections.defaultdict(collections.Counter); static = collections.Counter()
base = None
for r in rows[hdr[which] + 1:end]:
    if len(r) != len(H): continue
    a = int(r[ai], 16)
    if base is None: base = a
    k = region(amap.get(a - base))
    inst[k] += float(r[ii] or 0); smp[k] += float(r[si] or 0); static[k] += 1
    for h in stalls: st[k][h] += float(r[H.index(h)] or 0)
ti, ts = sum(inst.values()), sum(smp.values())
print(f"launch {which}: {ti:.3e} warp instructions, {ts:.0f} samples, {sum(static.values())} SASS instructions")
print(f"{'region':24s} {'static':>7s} {'inst%':>6s} {'smp%':>6s}  top stalls (share of the region's samples)")
for k, v in sorted(smp.items(), key=lambda kv: -kv[1]):
    top = sorted(st[k].items(), key=lambda kv: -kv[1])[:4]
    tot = max(1.0, sum(st[k].values()))
    print(f"{k:24s} {static[k]:7d} {inst[k] / ti * 100:6.1f} {v / ts * 100:6.1f}  " + ", ".join(f"{h[6:]} {x / tot * 100:.0f}%" for h, x in top))
