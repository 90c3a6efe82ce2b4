#!/bin/bash
# synchronous mcv_rig_process call: frames per chunk x slots
run() { env $1 timeout 200 python bench.py --steps 16 --no-cpu-baseline --no-matching --no-sweep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$1] device', round(d['value']), 'e2e', round(d['e2e']['value']), 'sync', round(d['e2e']['sync_call_value']))"; }
run "MCV_RIG_CHUNK=16"
run "MCV_RIG_CHUNK=22"
run "MCV_RIG_CHUNK=32"
run "MCV_RIG_CHUNK=43"
run "MCV_RIG_CHUNK=64"
run "MCV_RIG_CHUNK=32 MCV_RIG_SLOTS=4"
run "MCV_RIG_CHUNK=26 MCV_RIG_SLOTS=5"
