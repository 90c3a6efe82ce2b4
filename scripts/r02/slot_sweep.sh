#!/bin/bash
# synchronous mcv_rig_process (128 frames, host buffers) vs slots in rotation and chunk size
for s in 3 4 6; do for c in 8 16 22 32; do
  MCV_RIG_SLOTS=$s python bench.py --steps 20 --chunk $c --no-cpu-baseline --no-sweep --no-matching 2>/dev/null > /tmp/cs.json
  python - "$s" "$c" <<'PY'
import json, sys
a = json.load(open("/tmp/cs.json"))
print("slots", sys.argv[1], "chunk", sys.argv[2], "sync_call", round(a["e2e"]["sync_call_value"]), "pipelined_e2e", round(a["e2e"]["value"]), "device", round(a["value"]))
PY
done; done
