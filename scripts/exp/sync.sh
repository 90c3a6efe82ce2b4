run() { env $1 python bench.py --steps 30 --no-cpu-baseline --no-matching 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), 'e2e', round(d['e2e']['value']), 'sync', round(d['e2e']['sync_call_value']))"; }
run "MCV_RIG_CHUNK=32"
run "MCV_RIG_CHUNK=64"
run "MCV_RIG_CHUNK=43"
run "MCV_RIG_CHUNK=16"
run "MCV_RIG_CHUNK=64 MCV_RIG_STAGGER=1"
