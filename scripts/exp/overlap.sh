run() { env $1 python bench.py --steps 30 --no-cpu-baseline --no-matching 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['e2e']['value']), round(d['ms_per_step'],3))"; }
run "MCV_RIG_SLOTS_DEV=2"
run "MCV_RIG_SLOTS_DEV=2 MCV_RIG_QUAD_PRIORITY=1"
run "MCV_RIG_SLOTS_DEV=3 MCV_RIG_QUAD_PRIORITY=1"
run "MCV_RIG_SLOTS_DEV=1 MCV_RIG_QUAD_PRIORITY=1"
run "MCV_RIG_SLOTS_DEV=3 MCV_RIG_QUAD_PRIORITY=1 MCV_RIG_CHUNK_DEV=64"
