#!/bin/bash
# synchronous host-in/host-out call (mcv_rig_process) vs the engine's chunk size: 128 frames per call
for c in 4 8 16 32 64; do
  python bench.py --steps 20 --chunk $c --no-cpu-baseline --no-sweep --no-matching 2>/dev/null > /tmp/cs.json
  python - "$c" <<'PY'
import json, sys
a = json.load(open("/tmp/cs.json"))
print("chunk", sys.argv[1], "sync_call", round(a["e2e"]["sync_call_value"]), "pipelined_e2e", round(a["e2e"]["value"]), "device", round(a["value"]))
PY
done
