#!/usr/bin/env python
"""Per-stage DRAM traffic of one step from an ncu --set full report of every kernel (scripts/r04/final_profiles.sh): the first
captured launch of each kernel (all seven k_resize_march launches), dram__bytes_read.sum + dram__bytes_write.sum, per image.
    python scripts/ncu_traffic.py gpurun_out/r04f/all_full.ncu-rep 384 > profiles/r04_ncu_traffic.json"""
import csv, io, json, subprocess, sys
rep, n_images = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 384
STAGE = {"k_copy_level0": "pyramid", "k_gray_level0": "pyramid", "k_resize_march": "pyramid", "k_resize_level": "pyramid", "k_gauss7": "blur",
         "k_fast_score": "fast_score", "k_nms_sparse": "nms_cells", "k_cell_order": "nms_cells", "k_cell_fallback": "nms_cells",
         "k_octree_prep": "quadtree", "k_octree_replay": "quadtree", "k_octree": "quadtree", "k_orient_desc": "orient_desc",
         "k_stereo_rows": "stereo_match", "k_stereo_match": "stereo_match", "k_stereo_median": "stereo_median", "k_fill_tails": "stereo_median"}
LIMIT = {"k_resize_march": 7}
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
def val(r, name):
    return float(r[col[name]].replace(",", "")) * scale.get(units[col[name]], 1.0)
seen, per_image, us = {}, {}, {}
for r in data:
    name = r[col["Kernel Name"]]
    key = next((k for k in STAGE if k in name), None)
    if key is None:
        continue
    seen[key] = seen.get(key, 0) + 1
    if seen[key] > LIMIT.get(key, 1):
        continue
    st = STAGE[key]
    b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    per_image[st] = per_image.get(st, 0.0) + b / n_images
    us[st] = us.get(st, 0.0) + val(r, "gpu__time_duration.sum")
out = {"source": "ncu --set full --clock-control none: every kernel of a 128-frame step, first captured launch of each (%s, scripts/r04/final_profiles.sh), %d images per launch" % (rep, n_images),
       "unit": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per 640x480 image, per kernel launch set of the stage",
       "per_image": {k: int(round(v)) for k, v in per_image.items()},
       "kernel_us_serialised": {k: round(v, 1) for k, v in us.items()},
       "note": "stereo stages are per frame / 3; writes are under-counted where the L2 still holds dirty lines at kernel end; ncu times are cold-cache and serialised",
       "total_per_128_frame_step_GB": round(sum(per_image.values()) * n_images / 1e9, 3)}
print(json.dumps(out, indent=1))
