#!/usr/bin/env python
"""The reference's SHIPPED configuration (config/extractor.yaml: nkeypoints 2000, nlevels 1; 512 x 512 cameras): device-resident
throughput and one-triplet latency of the rig, with per-stage times."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mcvslam_b200.api as A
from mcvslam_b200 import synth
dev = torch.device("cuda", 0)
W = H = 512
for nlevels in (1, 8):
    for nb in (1, 16, 128):
        s = torch.cuda.Stream(device=dev)
        rig = A.Rig(nkeypoints=2000, scale_factor=1.2, nlevels=nlevels, device=0, stream=s.cuda_stream)
        cap = rig.cap
        base = [synth.triplet(50 + i, W, H) for i in range(min(nb, 16))]
        fr = torch.from_numpy(np.stack([base[i % len(base)] for i in range(nb)])).to(dev)
        sets = [(torch.empty(nb * 3 * cap * 28, dtype=torch.uint8, device=dev), torch.empty(nb * 3 * cap * 32, dtype=torch.uint8, device=dev),
                 torch.zeros(nb * 3, dtype=torch.int32, device=dev), torch.empty(nb * cap, dtype=torch.float32, device=dev),
                 torch.empty(nb * cap, dtype=torch.float32, device=dev)) for _ in range(6)]     # calls in flight never share outputs
        k, d, c, u, z = sets[0]
        with torch.cuda.stream(s):
            for _ in range(5):
                rig.process_async(fr.data_ptr(), nb, W, H, k.data_ptr(), d.data_ptr(), c.data_ptr(), u.data_ptr(), z.data_ptr()); rig.join()
            torch.cuda.synchronize()
            reps = 30
            t0 = time.perf_counter()
            for i in range(reps):
                rig.process_async(fr.data_ptr(), nb, W, H, *(t.data_ptr() for t in sets[i % 6]))
            rig.join(); torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            rig.set_profiling(True)
            for _ in range(10):
                rig.process_async(fr.data_ptr(), nb, W, H, k.data_ptr(), d.data_ptr(), c.data_ptr(), u.data_ptr(), z.data_ptr()); rig.join()
                torch.cuda.synchronize()
            ms, n = rig.stage_ms()
        print("nlevels %d, %3d frames/call: %.3f ms per call = %.0f frames/s; keypoints/img %.0f; stages us:" % (nlevels, nb, dt * 1e3, nb / dt, c.float().mean().item()),
              {a: round(1e3 * b / n, 1) for a, b in ms.items()})
