#!/usr/bin/env python
"""Per-stage CUDA-event times of ONE three-camera frame per call (the reference's Frame-per-call pattern), device-resident."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mcvslam_b200.api as A
from mcvslam_b200 import synth
dev = torch.device("cuda", 0)
for nb in (1, 4, 16):
    s = torch.cuda.Stream(device=dev)
    rig = A.Rig(device=0, stream=s.cuda_stream)
    cap = rig.cap
    fr = torch.from_numpy(np.stack([synth.triplet(50 + i) for i in range(nb)])).to(dev)
    k = torch.empty(nb * 3 * cap * 28, dtype=torch.uint8, device=dev); d = torch.empty(nb * 3 * cap * 32, dtype=torch.uint8, device=dev)
    c = torch.zeros(nb * 3, dtype=torch.int32, device=dev); u = torch.empty(nb * cap, dtype=torch.float32, device=dev); z = torch.empty(nb * cap, dtype=torch.float32, device=dev)
    with torch.cuda.stream(s):
        for _ in range(5):
            rig.process_async(fr.data_ptr(), nb, 640, 480, k.data_ptr(), d.data_ptr(), c.data_ptr(), u.data_ptr(), z.data_ptr()); rig.join()
        torch.cuda.synchronize()
        rig.set_profiling(True)
        for _ in range(20):
            rig.process_async(fr.data_ptr(), nb, 640, 480, k.data_ptr(), d.data_ptr(), c.data_ptr(), u.data_ptr(), z.data_ptr()); rig.join()
            torch.cuda.synchronize()
        ms, n = rig.stage_ms()
    print(nb, "frames/call:", {a: round(1e3 * b / n, 1) for a, b in ms.items()}, "us; sum", round(1e3 * sum(ms.values()) / n, 1))
