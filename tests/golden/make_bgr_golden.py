"""Generates tests/golden/bgr_golden.npz with the REAL cv2.cvtColor(COLOR_BGR2GRAY) (cv2 4.13 in the build container): the
fixture that pins the oracle's / the kernel's fixed-point BGR -> gray conversion (System::Track, src/System.cpp:60-64).
    python tests/golden/make_bgr_golden.py"""
import os

import cv2
import numpy as np

rng = np.random.default_rng(20261017)
out = {}
# 1. random image with an odd width (rows not 4-byte aligned in the 3-channel layout)
a = rng.integers(0, 256, (61, 97, 3), dtype=np.uint8)
# 2. channel sweep: every (B, G, R) with two channels on a coarse lattice and one channel full range
lat = np.array([0, 1, 2, 63, 127, 128, 129, 200, 254, 255], np.uint8)
full = np.arange(256, dtype=np.uint8)
sweeps = []
for ch in range(3):
    g = np.stack(np.meshgrid(full, lat, lat, indexing="ij"), -1).reshape(-1, 3)
    sweeps.append(np.roll(g, ch, axis=1))
b = np.concatenate(sweeps).reshape(-1, 256, 3).astype(np.uint8)
# 3. a rendered scene: three differently tinted copies of a synthetic image
yy, xx = np.mgrid[0:120, 0:160]
base = ((xx * 3 + yy * 5) % 256).astype(np.uint8)
c = np.stack([base, np.roll(base, 7, 1), 255 - base], -1)
for name, img in (("rand", a), ("sweep", b), ("scene", c)):
    img = np.ascontiguousarray(img)
    out[name + "_bgr"] = img
    out[name + "_gray"] = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
out["cv2_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bgr_golden.npz"), **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
