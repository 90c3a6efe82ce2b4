#!/usr/bin/env python
"""Runs bench.py with the given extra args and prints a compact summary (development helper)."""
import json
import subprocess
import sys

out = subprocess.run([sys.executable, "bench.py"] + sys.argv[1:], capture_output=True, text=True)
lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not lines:
    print("bench failed:", out.stdout[-2000:], out.stderr[-3000:])
    sys.exit(1)
d = json.loads(lines[-1])
print(" ".join(sys.argv[1:]), "| value %.0f e2e %.0f ms/step %.3f launches %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("gpu_launches")))
if "stage_ms_per_step" in d:
    print("   stages:", {k: round(v, 3) for k, v in d["stage_ms_per_step"].items()}, "sum %.3f" % sum(d["stage_ms_per_step"].values()))
print("   clocks:", d.get("clocks"), "roofline:", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac")} if "roofline" in d else None)
