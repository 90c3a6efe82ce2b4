run() { env $1 python bench.py --steps 40 --no-cpu-baseline --no-matching $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['value']), 'e2e', round(d['e2e']['value']))"; }
run "MCV_X=1" "--inflight 2"
run "MCV_X=1" "--inflight 3"
run "MCV_X=1" "--inflight 4"
run "MCV_RIG_SUBMIT_CHUNK=128" "--inflight 3"
run "MCV_RIG_SUBMIT_CHUNK=128" "--inflight 4"
run "MCV_RIG_SUBMIT_CHUNK=128 MCV_RIG_SLOTS=2" "--inflight 3"
run "MCV_RIG_SUBMIT_CHUNK=32" "--inflight 3"
