#!/usr/bin/env python
"""SASS digest of mcvslam_b200/libmcv_b200.so: per kernel, how often the mnemonics occur that prove the hardware path (TMA,
tcgen05 / TMEM, mbarrier, byte SIMD, popc, ...). No GPU needed.   python scripts/sass_digest.py > profiles/rNN_sass_digest.md"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "mcvslam_b200", "libmcv_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
COLS = ["UTMALDG", "UTCIMMA", "UTCBAR", "LDTM", "SYNCS", "VABSDIFF4", "VIMNMX3", "VIMNMX", "IDP.4A", "POPC", "SHFL", "VOTE", "ATOMG", "ATOMS", "REDUX", "STL", "LDL"]
rows, cur, arch = [], None, set()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = [re.sub(r"^_ZN3mcv\d+", "", m.group(1))[:40], {c: 0 for c in COLS}, 0]
        rows.append(cur)
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        cur[2] += 1
        for c in COLS:
            if op == c or op.startswith(c + ".") or (c == "VIMNMX" and op.startswith("VIMNMX") and not op.startswith("VIMNMX3")):
                if c == "VIMNMX" and op.startswith("VIMNMX3"):
                    continue
                cur[1][c] += 1
                break
print("# SASS digest of mcvslam_b200/libmcv_b200.so (%s cubins; `cuobjdump -sass`, scripts/sass_digest.py)" % ", ".join(sorted(arch)))
print("# per kernel: instruction counts of the mnemonics that prove the hardware path (TMA, tcgen05 / TMEM, mbarrier, byte SIMD, popc)\n")
print("| kernel | SASS instr | " + " | ".join(COLS) + " |")
print("|---|---|" + "---|" * len(COLS))
tot = {c: 0 for c in COLS}
for name, cnt, n in rows:
    print("| `%s` | %d | " % (name, n) + " | ".join(str(cnt[c]) if cnt[c] else "" for c in COLS) + " |")
    for c in COLS:
        tot[c] += cnt[c]
print("| **total** | %d | " % sum(r[2] for r in rows) + " | ".join(str(tot[c]) for c in COLS) + " |")
