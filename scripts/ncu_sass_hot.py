#!/usr/bin/env python
"""Hot-spot view of one kernel from an ncu --set full report (SASS page): executed warp-instructions and stall samples
grouped in runs of consecutive instructions.  python scripts/ncu_sass_hot.py report.ncu-rep kernel_regex [chunk]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# first kernel instance only
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hdr_i[0]]
end = hdr_i[1] - 1 if len(hdr_i) > 1 else len(rows)
data = [r for r in rows[hdr_i[0] + 1:end] if len(r) == len(h)]
si, ii, ss = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
tot_i = sum(int(r[ii]) for r in data); tot_s = sum(int(r[ss]) for r in data)
print("kernel %s: %d SASS lines, %.1f M warp-inst, %d samples" % (rows[hdr_i[0] - 1][1][:60], len(data), tot_i / 1e6, tot_s))
for c0 in range(0, len(data), chunk):
    blk = data[c0:c0 + chunk]
    bi = sum(int(r[ii]) for r in blk); bs = sum(int(r[ss]) for r in blk)
    ops = {}
    for r in blk:
        op = r[si].split()[0] if not r[si].strip().startswith("@") else r[si].split()[1]
        ops[op.split(".")[0]] = ops.get(op.split(".")[0], 0) + int(r[ii])
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
    print("%5d-%5d inst %5.1f%% samples %5.1f%%  %s" % (c0, c0 + len(blk), 100.0 * bi / tot_i, 100.0 * bs / max(1, tot_s),
                                                      " ".join("%s:%.1f" % (k, v / 1e6) for k, v in top)))
