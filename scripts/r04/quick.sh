#!/bin/bash
# Quick check of a kernel change on ONE B200: the extraction / rig parity tests, then a short bench run with per-stage times.
O=gpurun_out/r04; mkdir -p $O
T=${1:-q}
timeout 600 python -m pytest tests -m gpu -q -x > $O/${T}_tests.log 2>&1; tail -n 3 $O/${T}_tests.log
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-matching --no-sweep > $O/${T}_bench.json 2> $O/${T}_bench.err
python - <<PY
import json
d=json.loads(open("$O/${T}_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
print({k: round(v,4) for k,v in d["stage_ms_per_step"].items()})
PY
