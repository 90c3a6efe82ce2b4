run() { MCV_NVCC_EXTRA="$1" python -m mcvslam_b200.build --force > /dev/null 2>&1; timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-matching $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']; print('$1 $2', round(d['value']), 'e2e', round(d['e2e']['value']), 'quadtree', round(s['quadtree'],3))"; }
run "-DMCV_OC_THREADS=256" ""
run "-DMCV_OC_THREADS=64" ""
run "-DMCV_OR_MINB=40" ""
run "-DMCV_OR_MINB=48" ""
run "-DMCV_OR_MINB=64" ""
