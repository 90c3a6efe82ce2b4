#!/usr/bin/env python
"""Drives every matching kernel once at the sizes VERDICT r01 #6 names, for an `ncu --set full -k regex:...` capture:
k_knn2_bf (2000 x 2000 and a configs[4] shard 131072 x 2^20), k_knn2_candidates, k_project_match (10 k MapPoints),
k_fuse_match, k_wnd_track, k_stereo_match (through the rig). Not a benchmark: numbers printed under ncu are never bench values."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import mcvslam_b200.api as A
from mcvslam_b200 import synth

big = "--big" in sys.argv
L = A.lib()
q = synth.descriptors(2000, 1); t = synth.descriptors(2000, 2)
A.Matcher.KnnMatch(q, t)
if big:
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu"); g.manual_seed(5)
    nq, nt = 131072, 1 << 20
    dq = torch.randint(0, 256, (nq, 32), dtype=torch.uint8, generator=g).to(dev)
    dt = torch.randint(0, 256, (nt, 32), dtype=torch.uint8, generator=g).to(dev)
    idx = torch.empty((nq, 2), dtype=torch.int32, device=dev); dst = torch.empty((nq, 2), dtype=torch.int32, device=dev)
    A._check(L.mcv_knn2_bf_device(dq.data_ptr(), nq, dt.data_ptr(), nt, 0, idx.data_ptr(), dst.data_ptr(), 0))
    torch.cuda.synchronize()
rng = np.random.default_rng(3)
lens = rng.integers(0, 70, 2000)
off = np.zeros(2001, np.int32); off[1:] = np.cumsum(lens)
cidx = rng.integers(0, 2000, off[-1]).astype(np.int32)
A.Matcher.KnnMatchCandidates(q, t, off, cidx)
# projection: 10 k MapPoints on one extracted image (configs[2])
img = synth.scene(55)
E = A.ORB(2000, 1.2, 8, 28, 15)
n, k, d = E.Extract(img)
n_mp = 10000
fx = fy = 955.40503 * 640 / 512; cx, cy = 320.0, 240.0
R = np.eye(3, dtype=np.float32); tt = np.array([0.02, -0.01, 0.03], np.float32)
src = rng.integers(0, n, n_mp)
z = rng.uniform(2, 50, n_mp).astype(np.float32)
u = k["x"][src] + rng.normal(0, 2.0, n_mp).astype(np.float32); v = k["y"][src] + rng.normal(0, 2.0, n_mp).astype(np.float32)
pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1).astype(np.float32)
pw = (pc - tt).astype(np.float32)
md = d[src].copy()
lvl = k["octave"][src].astype(np.int32)
for r_th in (5.0, 10.0):
    print("project", A.ProjectBunchMapPoints(k, d, 640, 480, E.mvScaleFactor, R, tt, [fx, fy, cx, cy], pw, md, lvl, r_th)[0])
Ow = (-R.T @ tt).astype(np.float32)
view = pw - Ow
nrm = (view / np.linalg.norm(view, axis=1, keepdims=True)).astype(np.float32)
dl = np.full(n, -1, np.float32)
print("fuse", A.FuseMatch(k, d, 640, 480, E.mvLevelSigma2, E.mvInvLevelSigma2, R, tt, Ow, [fx, fy, cx, cy], dl, 955.40503, pw, nrm, md, lvl)[0])
b = np.roll(img, (3, -5), (0, 1))
n2, k2, d2 = E.Extract(b)
qi = np.sort(rng.choice(n, 1200, replace=False)).astype(np.int32)
print("wnd", A.WndTrack(k, d, qi, k2, d2, 640, 480)[0])
frames = np.stack([synth.triplet(s) for s in range(40, 48)])
rig = A.Rig()
out = rig.process(frames)
print("rig", out["counts"].sum(), int((out["u_right"] >= 0).sum()))
