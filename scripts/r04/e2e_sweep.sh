#!/bin/bash
# mcv_rig_submit: steps in flight x slots (streams) x frames per chunk; prints device-resident and host-in/host-out frames/s
run() { env $1 timeout 200 python bench.py --steps 24 --no-cpu-baseline --no-matching --no-sweep --inflight $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$1] inflight $2 device', round(d['value']), 'e2e', round(d['e2e']['value']), 'sync', round(d['e2e']['sync_call_value']), 'ceiling', round(d['e2e']['copy_ceiling']['frames_per_s']))"; }
run "MCV_RIG_SLOTS_SUBMIT=4" 4
run "MCV_RIG_SLOTS_SUBMIT=5" 5
run "MCV_RIG_SLOTS_SUBMIT=6" 6
run "MCV_RIG_SLOTS_SUBMIT=6" 8
run "MCV_RIG_SLOTS_SUBMIT=3" 6
run "MCV_RIG_SUBMIT_CHUNK=64 MCV_RIG_SLOTS_SUBMIT=6" 3
run "MCV_RIG_SUBMIT_CHUNK=64 MCV_RIG_SLOTS_SUBMIT=6" 4
run "MCV_RIG_SUBMIT_CHUNK=32 MCV_RIG_SLOTS_SUBMIT=6" 2
