#!/usr/bin/env python
"""One kernel of an ncu --set full report: stall ratios, pipe utilisation, memory-path figures.
    python scripts/ncu_kernel_report.py report.ncu-rep kernel_regex"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, d = rows[0], rows[1], rows[2]
st = [(float(d[i]), n) for i, n in enumerate(h) if ("issue_stalled" in n and "per_issue_active.ratio" in n)]
for v, n in sorted(st, reverse=True)[:7]:
    print("   stall %.2f %s" % (v, n.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
for n in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__registers_per_thread", "launch__grid_size"]:
    if n in h:
        print("   %s = %s %s" % (n, d[h.index(n)], u[h.index(n)]))
