run() { MCV_NVCC_EXTRA="$1" python -m mcvslam_b200.build --force > /dev/null 2>&1; timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-matching $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['value']), round(d['e2e']['value']), d['stage_ms_per_step']['orient_desc'])"; }
run "-DDESC_WAVES=4" ""
run "-DDESC_WAVES=2" ""
run "-DDESC_WAVES=8" ""
run "-DDESC_WAVES=4" "--chunk 64"
run "-DDESC_WAVES=4" "--chunk 128"
