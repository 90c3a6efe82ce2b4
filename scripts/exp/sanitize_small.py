import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import mcvslam_b200.api as A
from mcvslam_b200 import synth
# single image (per-image API), odd geometry, seeds; then a 2-frame rig batch; LK; distinctive
E = A.ORB(1200, 1.2, 8, 28, 15)
n, k, d = E.Extract(synth.scene(5, 641, 479))
print("extract", n)
E2 = A.ORB(500, 1.2, 8, 20, 7)
n2, k2, d2 = E2.Extract(synth.scene(6, 330, 250))
print("extract small", n2)
rig = A.Rig(device=0)
out = rig.process(np.stack([synth.triplet(7), synth.triplet(8)]))
print("rig", out["counts"].ravel())
a = synth.scene(40); b = synth.shifted(a, 1.3, -0.7, 1)
pts = np.stack([np.linspace(-5, 645, 300), np.linspace(-5, 485, 300)], 1).astype(np.float32)
o, s, e = A.LkTrack(a, b, pts)
print("lk", int(s.sum()))
bi, bm, od = A.ComputeDistinctiveDescriptors(d[:50], [0, 20, 50])
print("distinctive", bi)
# matcher family: integer-pipe kernel (small), tensor-core kernel (>= 2^20 pairs, split + merge), batched image pairs
q = synth.descriptors(300, 1); t = synth.descriptors(500, 2)
print("knn popc", int(A.Matcher.KnnMatch(q, t).knn["distance"].min()))
q = synth.descriptors(700, 7); t = synth.descriptors(1300, 8)          # one-launch warp-per-query kernel (k_knn2_wq), ragged last tile
print("knn one-launch", int(A.Matcher.KnnMatch(q, t).knn["distance"].min()))
# the reference's shipped single-level configuration: fewer candidates than quota, equal-key drain by lanes
E1 = A.ORB(2000, 1.2, 1, 28, 15)
n1, k1, d1 = E1.Extract(synth.scene(9, 512, 512))
print("extract shipped config", n1)
from test_oracle_golden import _voc_fixture
voc, vdesc, vexp = _voc_fixture()
print("bow real vocabulary", len(A.Vocabulary(voc).transform(vdesc, 4)["bow_ids"]))
q = synth.descriptors(1000, 3); t = synth.descriptors(9000, 4)
print("knn tensor", int(A.Matcher.KnnMatch(q, t).knn["distance"].min()))
import torch
dev = torch.device("cuda", 0)
desc = torch.from_numpy(synth.descriptors(3 * 1100, 5).reshape(3, 1100, 32)).to(dev); cnt = torch.tensor([1100, 900, 1000], dtype=torch.int32, device=dev)
pq = torch.tensor([0, 1], dtype=torch.int32, device=dev); pt = torch.tensor([1, 2], dtype=torch.int32, device=dev)
idx = torch.zeros((2, 1100, 2), dtype=torch.int32, device=dev); dst = torch.zeros_like(idx)
A._check(A.lib().mcv_knn2_pairs_device(desc.data_ptr(), cnt.data_ptr(), 3, 1100, pq.data_ptr(), pt.data_ptr(), 2, idx.data_ptr(), dst.data_ptr(), 0))
torch.cuda.synchronize()
print("knn pairs", int(dst[0, :1100, 0].min()))
