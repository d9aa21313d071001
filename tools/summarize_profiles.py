#!/usr/bin/env python
"""Turn gpurun_out/{launches,prof_*}_<tag> into the committed summaries under profiles/ (run in the build container)."""
import collections, csv, io, os, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__t_bytes.sum", "smsp__thread_inst_executed.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
out = [f"# ncu summary, tag {tag}", ""]
lp = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[h]; ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[h + 1:]:
        if len(r) > vi:
            agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")) / (1e3 if r[ui] == "ns" else 1))
    tot = sum(sum(v) for v in agg.values())
    out += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised: compare SHARES)", "",
            "| kernel | launches | mean us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v) / tot * 100:.1f} % |")
    out.append("")
    with open(os.path.join(P, f"launches_{tag}.csv"), "w") as f:
        f.write(open(lp).read())
for name in sorted(os.listdir(G)):
    if not (name.startswith("prof_") and name.endswith(f"_{tag}.ncu-rep")):
        continue
    raw = subprocess.run(["ncu", "-i", os.path.join(G, name), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    H, U = rows[0], rows[1]
    for V in rows[2:]:
        kn = V[H.index("Kernel Name")] if "Kernel Name" in H else name
        out += [f"## `{name}` — {kn} (`ncu --set full --clock-control none`)", "", "| metric | value | unit |", "|---|---|---|"]
        for i, hname in enumerate(H):
            if hname in KEEP or ("issue_stalled" in hname and hname.endswith("_per_warp_active.pct")):
                out.append(f"| {hname} | {V[i]} | {U[i]} |")
        out.append("")
open(os.path.join(P, f"ncu_summary_{tag}.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:8000])
