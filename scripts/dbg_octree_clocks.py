import sys, numpy as np
sys.path.insert(0, ".")
import mcvslam_b200.api as A
from mcvslam_b200 import synth
imgs = np.stack([synth.scene(1000 + s) for s in range(192)])
E = A.ORB(2000, 1.2, 8, 28, 15)
for it in range(2):
    E.ExtractBatch(imgs)
    c = A.octree_clocks()
    d = np.diff(c[:6])
    print("phases (cycles): gather %d roots %d split %d drain %d select %d total %d" % (*d, c[5] - c[0]))
