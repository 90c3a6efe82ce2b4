#!/usr/bin/env python
"""Summarises an ncu --set full report as a markdown table (one row per profiled launch).
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.md"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time us", 1), ("dram__bytes_read.sum", "DRAM rd MB", 1), ("dram__bytes_write.sum", "DRAM wr MB", 1),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %", 1),
        ("launch__registers_per_thread", "regs", 0), ("smsp__inst_executed.sum", "warp-inst M", 1e-6), ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", 1)]


def to_unit(v, unit, want):
    v = float(v.replace(",", ""))
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
    return v * scale.get(unit, 1)


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki, gi, bi = hdr.index("Kernel Name"), hdr.index("launch__grid_size"), hdr.index("launch__block_size")
    print("| kernel | grid x block | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for r in data:
        cells = []
        for name, label, mul in COLS:
            i = hdr.index(name)
            v = to_unit(r[i], units[i], label)
            v = v * mul if mul not in (0, 1) else v
            cells.append("%d" % v if mul == 0 else ("%.1f" % v))
        print("| `%s` | %s x %s | %s |" % (r[ki].split("(")[0], r[gi], r[bi], " | ".join(cells)))


if __name__ == "__main__":
    main(sys.argv[1])
