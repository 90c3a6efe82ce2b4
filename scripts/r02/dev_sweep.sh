#!/bin/bash
# device-resident path: frames per internal chunk x slots (streams) in rotation
for s in 2 3 4 6; do for c in 32 64 128; do
  MCV_RIG_SLOTS_DEV=$s MCV_RIG_CHUNK_DEV=$c python bench.py --steps 30 --no-cpu-baseline --no-sweep --no-matching 2>/dev/null > /tmp/ds.json
  python - "$s" "$c" <<'PY'
import json, sys
a = json.load(open("/tmp/ds.json"))
print("slots_dev", sys.argv[1], "chunk_dev", sys.argv[2], "device", round(a["value"]), "e2e", round(a["e2e"]["value"]))
PY
done; done
