#!/usr/bin/env python
"""Pipe utilisation, memory-path throughput and top stall reasons per distinct launch of an ncu --set full report (markdown).
    python scripts/ncu_pipes.py report.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h, u, data = rows[0], rows[1], rows[2:]
ki = h.index("Kernel Name")
cols = [("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__warps_active.avg.per_cycle_active", "warps/SM"), ("smsp__warps_eligible.avg.per_cycle_active", "eligible/sched")]
print("| kernel (grid) | " + " | ".join(c[1] for c in cols) + " | top stalls (warps per issue) |")
print("|---|" + "---|" * (len(cols) + 1))
seen = set()
for d in data:
    name = d[ki].split("(")[0].replace("mcv::", "")
    g = d[h.index("launch__grid_size")]
    if (name, g) in seen:
        continue
    seen.add((name, g))
    st = sorted([(float(d[i]), n.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")) for i, n in enumerate(h)
                 if "issue_stalled" in n and "per_issue_active.ratio" in n and "selected" not in n], reverse=True)[:3]
    print("| `%s` (%s) | " % (name, g) + " | ".join("%.1f" % float(d[h.index(c[0])]) for c in cols) + " | " + ", ".join("%s %.2f" % (n, v) for v, n in st) + " |")
