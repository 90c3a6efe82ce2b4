run() { env $1 python bench.py --steps 40 --no-cpu-baseline --no-matching $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['value']), 'e2e', round(d['e2e']['value']))"; }
run "MCV_RIG_STAGGER=1" ""
run "MCV_RIG_STAGGER=0" ""
run "MCV_RIG_STAGGER=0 MCV_RIG_SLOTS_DEV=3" ""
run "MCV_RIG_STAGGER=0 MCV_RIG_SLOTS_DEV=3 MCV_RIG_CHUNK_DEV=64" ""
run "MCV_RIG_STAGGER=0 MCV_RIG_SLOTS_DEV=2" "--frames 256"
run "MCV_RIG_STAGGER=1 MCV_RIG_SLOTS_DEV=2" "--frames 256"
run "MCV_RIG_STAGGER=1 MCV_RIG_SLOTS_DEV=2" "--frames 64"
