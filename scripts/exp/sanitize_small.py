import sys
sys.path.insert(0, ".")
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
