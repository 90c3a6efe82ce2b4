#!/usr/bin/env python
"""One-image-pair brute-force matching (mcv_knn2_bf_device, device-resident): the one-launch warp-per-query kernel against the
split + merge pair it replaces at these sizes, and against the tensor-core path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench as B
import mcvslam_b200.api as A
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
peak, _ = A.popc_peak(8192)
for nq, nt in ((256, 2000), (500, 500), (1000, 1000), (2000, 2000), (2000, 4000), (500, 16000), (5000, 5000)):
    row = []
    for name, env in (("wq", {"MCV_KNN_POPC": "1"}), ("split+merge", {"MCV_KNN_POPC": "1", "MCV_KNN_NO_WQ": "1"}), ("shipped", {"MCV_KNN_POPC": "0"})):
        os.environ.pop("MCV_KNN_NO_WQ", None)
        os.environ.update(env)
        with torch.cuda.stream(stream):
            s = B.time_knn2(A, torch, dev, stream, nq, nt, 200, warm=5)
        row.append("%s %.1f us (%.0f %% of popc peak)" % (name, s * 1e6, 100 * nq * nt * 8 / s / peak))
    print(nq, "x", nt, ":", "; ".join(row))
