#!/usr/bin/env python
"""Phase clocks of the level-0 quadtree task of ONE image per call (k_octree_prep / k_octree_replay alone on the GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import mcvslam_b200.api as A
from mcvslam_b200 import synth
cfgs = [(8, 640, 480), (1, 512, 512)]      # BASELINE configs[1]; the reference's shipped config/extractor.yaml on its 512 x 512 cameras
for nlevels, W, H in cfgs:
  E = A.ORB(2000, 1.2, nlevels, 28, 15)
  for nb in (1, 3):
    imgs = np.stack([synth.scene(1000 + s, W, H) for s in range(nb)])
    for it in range(3):
        E.ExtractBatch(imgs)
        c = A.octree_clocks()
    d = np.diff(c[:6])
    print("%d level(s) %dx%d, images %d: prep gather %d sort %d | replay split %d drain %d select %d cycles" % (nlevels, W, H, nb, d[0], d[1], c[3] - c[6], d[3], d[4]))
