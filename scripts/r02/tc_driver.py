#!/usr/bin/env python
"""Drives the tensor-core matcher once per size for `ncu --set full -k regex:k_knn2_tc|k_expand_pm1`: 2000 x 2000 (configs[0]),
64 frames x 5000 consecutive pairs (configs[3], mcv_knn2_pairs_device) and a configs[4] shard (131072 x 2^20)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import mcvslam_b200.api as A
L = A.lib()
dev = torch.device("cuda", 0)
g = torch.Generator(device="cpu"); g.manual_seed(5)
s = torch.cuda.current_stream(dev).cuda_stream
def knn(nq, nt):
    q = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, generator=g).to(dev); t = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, generator=g).to(dev)
    idx = torch.empty((nq, 2), dtype=torch.int32, device=dev); dst = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    A._check(L.mcv_knn2_bf_device(q.data_ptr(), nq, t.data_ptr(), nt, 0, idx.data_ptr(), dst.data_ptr(), s)); torch.cuda.synchronize()
    return int(dst[:, 0].min())
print(knn(2000, 2000))
n_img, cap = 64, 5056
desc = torch.randint(0, 256, (n_img, cap, 32), dtype=torch.uint8, generator=g).to(dev); cnt = torch.full((n_img,), 5000, dtype=torch.int32, device=dev)
pq = torch.arange(0, n_img - 1, dtype=torch.int32, device=dev); pt = pq + 1
idx = torch.empty((n_img - 1, cap, 2), dtype=torch.int32, device=dev); dst = torch.empty_like(idx)
A._check(L.mcv_knn2_pairs_device(desc.data_ptr(), cnt.data_ptr(), n_img, cap, pq.data_ptr(), pt.data_ptr(), n_img - 1, idx.data_ptr(), dst.data_ptr(), s)); torch.cuda.synchronize()
if "--big" in sys.argv:
    print(knn(131072, 1 << 20))
