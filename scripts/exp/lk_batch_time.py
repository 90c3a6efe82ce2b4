import time, numpy as np, sys
sys.path.insert(0, ".")
import mcvslam_b200.api as A
from mcvslam_b200 import synth
NP = 64
a = np.stack([synth.scene(40 + (k % 4)) for k in range(NP)]); b = np.stack([synth.shifted(a[k], 1.3, -0.7, k) for k in range(NP)])
rng = np.random.default_rng(0)
pts = [np.stack([rng.uniform(20, 620, 2000), rng.uniform(20, 460, 2000)], 1).astype(np.float32) for _ in range(NP)]
A.LkTrackBatch(a, b, pts)
t = time.perf_counter()
for _ in range(10): A.LkTrackBatch(a, b, pts)
dt = (time.perf_counter() - t) / 10
print("LkTrackBatch %d pairs x 2000 pts 640x480 host-in/host-out: %.3f ms per call = %.1f us per pair" % (NP, dt * 1e3, dt * 1e6 / NP))
