run() { MCV_NVCC_EXTRA="$1" python -m mcvslam_b200.build --force > /dev/null 2>&1; timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-matching $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stage_ms_per_step']; print('$1 $2', round(d['value']), 'e2e', round(d['e2e']['value']), 'nms', round(s['nms_cells'],3), 'desc', round(s['orient_desc'],3))"; }
run "-DMCV_TMA_L2PROMO=CU_TENSOR_MAP_L2_PROMOTION_NONE" ""
run "-DMCV_TMA_L2PROMO=CU_TENSOR_MAP_L2_PROMOTION_L2_64B" ""
run "-DMCV_TMA_L2PROMO=CU_TENSOR_MAP_L2_PROMOTION_L2_256B" ""
