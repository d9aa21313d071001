#!/usr/bin/env python
"""Attribute an ncu source-page CSV of lidar_kernel to the blocks of its state machine (lidar.cu) and report, per block, the
warp instructions, thread instructions and the mean number of active lanes (SIMT utilisation), using nvdisasm -g line info
of the same cubin (inlined callees are attributed to the line of the kernel that called them).
usage: ncu_lidar_blocks.py <source_page.csv> <nvdisasm -g -c output> [kernel substring]"""
import csv, re, collections, sys, os
csvp, disp = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "lidar_kernelILb0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "ft_grandprix_b200", "csrc", "lidar.cu")).read().split("\n")
# block boundaries: the "// ----------------" banners inside the kernel, plus named sub-blocks
marks = []
for i, line in enumerate(src, 1):
    m = re.search(r"// ---------------- (.*)$", line) or re.search(r"//\s*\[block: (.*?)\]", line)
    if m: marks.append((i, m.group(1)[:44]))
kstart = next(i for i, l in enumerate(src, 1) if "lidar_kernel(const uint32_t*" in l)
kend = next(i for i, l in enumerate(src, 1) if l.startswith("static int ensure_beams"))
marks = [(kstart, "prologue / batch frames")] + [m for m in marks if kstart < m[0] < kend]
def block(line):
    if line is None or not (kstart <= line < kend): return "?"
    name = marks[0][1]
    for l, n in marks:
        if line >= l: name = n
    return name
amap = {}; cur = None; inside = False
for line in open(disp):
    if line.startswith(".text."):
        inside = kern in line
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', line)
    if m:
        f, l, rest = m.group(1).split("/")[-1], int(m.group(2)), m.group(3)
        # outermost call site inside lidar.cu's kernel
        sites = [(f, l)] + [(a.split("/")[-1], int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', rest)]
        cur = None
        for ff, ll in sites:
            if ff == "lidar.cu" and kstart <= ll < kend: cur = ll
        if cur is None: cur = -1
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", line)
    if m: amap[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(csvp)))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
H = rows[hdr]
ai, si, ii, ti = H.index("Address"), H.index("# Samples"), H.index("Instructions Executed"), H.index("Thread Instructions Executed")
inst = collections.Counter(); thr = collections.Counter(); smp = collections.Counter(); static = collections.Counter()
linst = collections.Counter(); lthr = collections.Counter()
base = None
for r in rows[hdr + 1:]:
    if len(r) != len(H): continue
    a = int(r[ai], 16)
    if base is None: base = a
    ln = amap.get(a - base)
    k = block(ln if ln and ln > 0 else None)
    inst[k] += float(r[ii] or 0); thr[k] += float(r[ti] or 0); smp[k] += float(r[si] or 0); static[k] += 1
    linst[ln] += float(r[ii] or 0); lthr[ln] += float(r[ti] or 0)
TI, TT, TS = sum(inst.values()), sum(thr.values()), sum(smp.values())
print(f"{TI:.3e} warp instructions, {TT:.3e} thread instructions, mean active lanes {TT / TI:.2f}, {TS:.0f} samples")
print(f"{'block':46s} {'static':>6s} {'winst%':>7s} {'tinst%':>7s} {'lanes':>6s} {'smp%':>6s}")
for k, v in sorted(inst.items(), key=lambda kv: -kv[1]):
    print(f"{k:46s} {static[k]:6d} {v / TI * 100:7.1f} {thr[k] / TT * 100:7.1f} {thr[k] / max(v, 1):6.1f} {smp[k] / max(TS, 1) * 100:6.1f}")
if "--lines" in sys.argv:
    print("\nhottest kernel lines (warp inst %, lanes)")
    for ln, v in sorted(linst.items(), key=lambda kv: -kv[1])[:400]:
        txt = src[ln - 1].strip()[:90] if ln and ln > 0 else "?"
        print(f"{str(ln):>5s} {v / TI * 100:5.1f} {lthr[ln] / max(v, 1):5.1f}  {txt}")
