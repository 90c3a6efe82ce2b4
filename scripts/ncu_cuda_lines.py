#!/usr/bin/env python
"""Per CUDA source line: executed warp-instructions and stall samples of one kernel from an ncu --set full report taken with
--import-source on.   python scripts/ncu_cuda_lines.py report.ncu-rep kernel_regex [top_n]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file, agg, hdr = None, {}, None
first_fn = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        if first_fn is None:
            first_fn = r[1]
        elif r[1] != first_fn:
            pass
    elif r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == "-":      # a CUDA source line row (aggregated over its SASS)
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()])
        a[0] += int(r[ii]); a[1] += int(r[si])
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total %.1f M warp-inst, %d samples (launch instances summed)" % (ti / 1e6, ts))
for (f, ln), (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100.0 * i / ti, 100.0 * s / max(ts, 1), f, ln, src[:110]))
