#!/bin/bash
# Build variants on the GPU box and print the per-stage times of a short bench run. usage: variants.sh "<nvcc extra 1>" "<nvcc extra 2>" ...
run() { MCV_NVCC_EXTRA="$1" python -m mcvslam_b200.build --force > /dev/null 2>&1; timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-matching --no-sweep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']; print('[$1]', round(d['value']), 'ms', round(d['ms_per_step'],4), {k: round(v,4) for k,v in s.items()})"; }
for v in "$@"; do run "$v"; done
