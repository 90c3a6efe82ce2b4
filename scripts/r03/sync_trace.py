#!/usr/bin/env python
"""Timeline of the chunks of one synchronous mcv_rig_process call on pinned host buffers (needs an experiment build:
MCV_NVCC_EXTRA=-DMCV_EXPERIMENTS python -m mcvslam_b200.build --force; MCV_RIG_TRACE=1)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench as B
import mcvslam_b200.api as A
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 128
frames = B.make_frames(nb, 100)
rig = A.Rig(device=0)
h = torch.from_numpy(frames).pin_memory()
cap = rig.cap
k = torch.empty(nb * 3 * cap * 28, dtype=torch.uint8).pin_memory(); d = torch.empty(nb * 3 * cap * 32, dtype=torch.uint8).pin_memory()
c = torch.zeros(nb * 3, dtype=torch.int32).pin_memory(); u = torch.empty(nb * cap, dtype=torch.float32).pin_memory(); z = torch.empty(nb * cap, dtype=torch.float32).pin_memory()
L = A.lib()
def call():
    A._check(L.mcv_rig_process(rig._r, h.data_ptr(), nb, 640, 480, 0, k.data_ptr(), d.data_ptr(), c.data_ptr(), u.data_ptr(), z.data_ptr(), cap, 0))
for _ in range(6):
    t0 = time.perf_counter(); call(); dt = time.perf_counter() - t0
    print("call %.3f ms" % (dt * 1e3), file=sys.stderr)
