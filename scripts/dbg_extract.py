import sys, numpy as np
sys.path.insert(0, ".")
import mcvslam_b200.api as A
from mcvslam_b200 import synth
from oracle import oracle as O
img = synth.scene(1234)
E = A.ORB(2000, 1.2, 8, 28, 15)
print("extract...", flush=True)
n, k, d = E.Extract(img)
print("n", n, flush=True)
Or = O.Orb(2000, 1.2, 8, 28, 15, debug=True)
no, ko, do = Or.extract(img)
print("oracle n", no)
for l in range(8):
    a = E.debug_level_keypoints(l, 0); b = Or.debug_kps(0, l)
    same = len(a) == len(b) and (a["x"] == b["x"]).all() and (a["y"] == b["y"]).all() and (a["response"] == b["response"]).all()
    print("level", l, len(a), len(b), "cand same", same, flush=True)
print("kps same", k.tobytes() == ko.tobytes(), "desc same", d.tobytes() == do.tobytes())
