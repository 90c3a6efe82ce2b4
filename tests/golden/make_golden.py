#!/usr/bin/env python
"""Generates tests/golden/*.npz — the fixtures that pin the oracle.

The reference (Sologala/MCVSLAM) ships no golden vectors for this path and cannot be built here.
What CAN run here is the third-party library that owns the arithmetic: OpenCV (Python cv2 4.13.0). This script
drives the REAL cv2 primitives the reference calls — cv2.resize, cv2.FastFeatureDetector (per cell, on the same
sub-images, with the ini->min threshold fallback), cv2.fastAtan2, cv2.GaussianBlur, cv2.BFMatcher.knnMatch — in the
order ORBextractor.cc:831-919 / Matcher.cpp call them, with an independent pure-Python restatement of the first-party
glue (cell grid ORBextractor.cc:582-633, quadtree :469-580 with a libstdc++ push_heap/pop_heap emulation, IC_Angle
:75-98, steered rBRIEF :101-141). Its outputs are committed; tests compare the C++ oracle and the CUDA path to them.

Run from the repo root in the build container:  python tests/golden/make_golden.py
"""
import math
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from mcvslam_b200 import synth  # noqa: E402

cv2.setNumThreads(1)
f32 = np.float32
HERE = os.path.dirname(os.path.abspath(__file__))


def load_pattern():
    txt = open(os.path.join(ROOT, "oracle", "ora_pattern.inc")).read()
    nums = [int(t) for line in txt.splitlines() if not line.startswith("//") for t in line.split(",") if t.strip()]
    return np.array(nums, np.int32).reshape(512, 2)


PATTERN = load_pattern()


def rint(v):  # cvRound: round-half-even
    return int(np.rint(v))


class Params:
    def __init__(self, nfeatures, scale_factor, nlevels, ini_th, min_th):
        self.nfeatures, self.nlevels, self.ini_th, self.min_th = nfeatures, nlevels, ini_th, min_th
        sf = f32(scale_factor)
        self.scale = [f32(1.0)]
        for i in range(1, nlevels):
            self.scale.append(f32(self.scale[-1] * sf))
        self.inv_scale = [f32(f32(1.0) / s) for s in self.scale]
        factor = f32(f32(1.0) / sf)
        nd = f32(f32(f32(nfeatures) * f32(f32(1) - factor)) / f32(f32(1) - f32(math.pow(float(factor), float(nlevels)))))
        self.quota = []
        tot = 0
        for _ in range(nlevels - 1):
            self.quota.append(rint(nd)); tot += self.quota[-1]; nd = f32(nd * factor)
        self.quota.append(max(nfeatures - tot, 0))
        umax = [0] * 16
        vmax = int(math.floor(float(f32(f32(15) * f32(math.sqrt(2.0)) / f32(2) + f32(1)))))
        vmin = int(math.ceil(float(f32(f32(15) * f32(math.sqrt(2.0)) / f32(2)))))
        for v in range(vmax + 1):
            umax[v] = rint(math.sqrt(225.0 - v * v))
        v0 = 0
        for v in range(15, vmin - 1, -1):
            while umax[v0] == umax[v0 + 1]:
                v0 += 1
            umax[v] = v0; v0 += 1
        self.umax = umax


def pyramid(img, P):
    lv = [img.copy()]
    h, w = img.shape
    for l in range(1, P.nlevels):
        dw, dh = rint(f32(w) * P.inv_scale[l]), rint(f32(h) * P.inv_scale[l])
        lv.append(cv2.resize(lv[-1], (dw, dh), interpolation=cv2.INTER_LINEAR))
    return lv


# --- libstdc++ heap emulation (bits/stl_heap.h __push_heap / __adjust_heap), comparator = size(a) < size(b) ---
def push_heap(h, key):
    hole = len(h) - 1
    val = h[hole]
    parent = (hole - 1) // 2
    while hole > 0 and key(h[parent]) < key(val):
        h[hole] = h[parent]; hole = parent; parent = (hole - 1) // 2
    h[hole] = val


def pop_heap(h, key):
    # moves the top to the back; caller pops it
    last = len(h) - 1
    val = h[last]
    h[last] = h[0]
    n = last
    hole = 0
    child = 0
    while child < (n - 1) // 2:
        child = 2 * (child + 1)
        if key(h[child]) < key(h[child - 1]):
            child -= 1
        h[hole] = h[child]; hole = child
    if (n & 1) == 0 and child == (n - 2) // 2:
        child = 2 * (child + 1)
        h[hole] = h[child - 1]; hole = child - 1
    parent = (hole - 1) // 2
    while hole > 0 and key(h[parent]) < key(val):
        h[hole] = h[parent]; hole = parent; parent = (hole - 1) // 2
    h[hole] = val


class Node:
    __slots__ = ("ulx", "uly", "urx", "bry", "keys")


def divide(n):
    halfx = int(math.ceil(float(f32(n.urx - n.ulx) / f32(2))))
    halfy = int(math.ceil(float(f32(n.bry - n.uly) / f32(2))))
    mx, my = n.ulx + halfx, n.uly + halfy
    c = []
    for (ulx, uly, urx, bry) in ((n.ulx, n.uly, mx, my), (mx, n.uly, n.urx, my), (n.ulx, my, mx, n.bry), (mx, my, n.urx, n.bry)):
        k = Node(); k.ulx, k.uly, k.urx, k.bry, k.keys = ulx, uly, urx, bry, []
        c.append(k)
    for kp in n.keys:
        if kp[0] < mx:
            (c[0] if kp[1] < my else c[2]).keys.append(kp)
        elif kp[1] < my:
            c[1].keys.append(kp)
        else:
            c[3].keys.append(kp)
    return [k for k in c if k.keys]


def distribute(cands, min_x, max_x, min_y, max_y, N):
    n_ini = int(np.round(f32(max_x - min_x) / f32(max_y - min_y)))  # std::round(float): half away; never a tie here
    hx = f32(f32(max_x - min_x) / f32(n_ini))
    roots = []
    for i in range(n_ini):
        k = Node(); k.ulx = int(f32(hx * f32(i))); k.urx = int(f32(hx * f32(i + 1))); k.uly = 0; k.bry = max_y - min_y; k.keys = []
        roots.append(k)
    for kp in cands:
        roots[int(f32(kp[0]) / hx)].keys.append(kp)
    key = lambda nd: len(nd.keys)
    heap = []
    for r in roots:
        if r.keys:
            heap.append(r); push_heap(heap, key)
    if not heap:
        return []
    while len(heap) < N:
        top = heap[0]
        if len(top.keys) == 1:
            break
        pop_heap(heap, key); heap.pop()
        for s in divide(top):
            heap.append(s); push_heap(heap, key)
    out = []
    while heap:
        top = heap[0]
        pop_heap(heap, key); heap.pop()
        best = top.keys[0]
        for kp in top.keys[1:]:
            if kp[2] > best[2]:
                best = kp
        out.append(best)
    return out


def detect_level(img, P, level):
    """Cell loop ORBextractor.cc:582-633 with real cv2.FAST on each cell sub-image. Returns (cands, kept)."""
    h, w = img.shape
    minb, maxbx, maxby = 16, w - 16, h - 16
    width, height = f32(maxbx - minb), f32(maxby - minb)
    ncols, nrows = int(width / f32(35)), int(height / f32(35))
    wcell, hcell = int(math.ceil(float(width / f32(ncols)))), int(math.ceil(float(height / f32(nrows))))
    det_ini = cv2.FastFeatureDetector_create(P.ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_min = cv2.FastFeatureDetector_create(P.min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    cands = []
    for i in range(nrows):
        iniy = minb + i * hcell; maxy = iniy + hcell + 6
        if iniy >= maxby - 3:
            continue
        maxy = min(maxy, maxby)
        for j in range(ncols):
            inix = minb + j * wcell; maxx = inix + wcell + 6
            if inix >= maxbx - 6:
                continue
            maxx = min(maxx, maxbx)
            cell = np.ascontiguousarray(img[iniy:maxy, inix:maxx])
            k = det_ini.detect(cell)
            if not k:
                k = det_min.detect(cell)
            for p in k:
                cands.append((int(p.pt[0]) + j * wcell, int(p.pt[1]) + i * hcell, int(p.response)))
    kept = distribute(cands, minb, maxbx, minb, maxby, P.quota[level])
    return cands, kept


def ic_angle(img, x, y, umax):
    p = img.astype(np.int64)
    m10 = sum(u * p[y, x + u] for u in range(-15, 16))
    m01 = 0
    for v in range(1, 16):
        d = umax[v]
        vs = 0
        for u in range(-d, d + 1):
            a, b = p[y + v, x + u], p[y - v, x + u]
            vs += a - b; m10 += u * (a + b)
        m01 += v * vs
    return f32(cv2.fastAtan2(float(f32(m01)), float(f32(m10))))


def sincosf_glibc(a):
    """glibc 2.39 sincosf restated in double (verified == libm over all 1.09e9 floats in [0, 6.5])."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    s, c = ctypes.c_float(), ctypes.c_float()
    libm.sincosf.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    libm.sincosf(ctypes.c_float(float(a)), ctypes.byref(s), ctypes.byref(c))
    return f32(s.value), f32(c.value)


def descriptor(blur, x, y, angle):
    ang = f32(f32(angle) * f32(np.pi / f32(180.0)))  # factorPI = (float)(CV_PI/180.f)
    b, a = sincosf_glibc(ang)
    px = PATTERN[:, 0].astype(f32); py = PATTERN[:, 1].astype(f32)
    dy = np.rint((px * b).astype(f32) + (py * a).astype(f32)).astype(np.int64)   # separate f32 mul/add
    dx = np.rint((px * a).astype(f32) - (py * b).astype(f32)).astype(np.int64)
    v = blur[y + dy, x + dx].astype(np.int32)
    bits = (v[0::2] < v[1::2]).astype(np.uint8).reshape(32, 8)
    return (bits << np.arange(8, dtype=np.uint8)).sum(axis=1).astype(np.uint8)


def extract(img, P):
    lv = pyramid(img, P)
    kps, descs, cand_counts = [], [], []
    for l, im in enumerate(lv):
        cands, kept = detect_level(im, P, l)
        cand_counts.append(len(cands))
        if not kept:
            continue
        blur = cv2.GaussianBlur(im.copy(), (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        size = f32(int(f32(f32(31) * P.scale[l])))
        for (cx, cy, resp) in kept:
            x, y = cx + 16, cy + 16
            ang = ic_angle(im, x, y, P.umax)
            descs.append(descriptor(blur, x, y, ang))
            fx, fy = f32(x), f32(y)
            if l != 0:
                fx, fy = f32(fx * P.scale[l]), f32(fy * P.scale[l])
            kps.append((fx, fy, size, ang, f32(resp), l, -1))
    kp_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                         ("octave", "<i4"), ("class_id", "<i4")])
    return np.array(kps, kp_dtype), np.array(descs, np.uint8).reshape(-1, 32), lv, np.array(cand_counts, np.int32)


def bf_knn2(q, t):
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    idx = np.array([[e.trainIdx for e in r] for r in m], np.int32)
    dist = np.array([[e.distance for e in r] for r in m], np.float32)
    return idx, dist


def main():
    out = {}
    # G1: 640x480, 8 levels (BASELINE config #1/#2 geometry), two images + BF match between them
    P = Params(2000, 1.2, 8, 28, 15)
    imgs = [synth.scene(1000), synth.scene(1001)]
    res = [extract(im, P) for im in imgs]
    for i, (k, d, lv, cc) in enumerate(res):
        out[f"g1_kps{i}"], out[f"g1_desc{i}"], out[f"g1_cand_counts{i}"] = k, d, cc
        out[f"g1_level7_{i}"] = lv[7]
    idx, dist = bf_knn2(res[0][1], res[1][1])
    out["g1_bf_idx"], out["g1_bf_dist"] = idx, dist
    # G2: shipped config (nlevels 1, config/extractor.yaml) on a 512x512 image (config/camleft.yaml cx=cy=256)
    P1 = Params(2000, 1.2, 1, 28, 15)
    k, d, lv, cc = extract(synth.scene(42, 512, 512), P1)
    out["g2_kps"], out["g2_desc"], out["g2_cand_counts"] = k, d, cc
    # G3: small odd-sized low-texture image so that the minTh fallback and quota shortfall are exercised
    P3 = Params(300, 1.2, 4, 28, 15)
    rng = np.random.default_rng(3)
    im3 = synth.scene(77, 217, 163)
    im3[:, :90] = (im3[:, :90].astype(np.int32) // 4 + 96).astype(np.uint8)  # flatten contrast on the left part
    k, d, lv, cc = extract(im3, P3)
    out["g3_img"], out["g3_kps"], out["g3_desc"], out["g3_cand_counts"] = im3, k, d, cc
    # G4: tie-heavy brute-force 2-NN (low-entropy descriptors)
    q = synth.descriptors(300, 11, low_entropy=True); t = synth.descriptors(500, 12, low_entropy=True)
    q[:, 4:] = 0; t[:, 4:] = 0
    idx, dist = bf_knn2(q, t)
    out["g4_q"], out["g4_t"], out["g4_idx"], out["g4_dist"] = q, t, idx, dist
    np.savez_compressed(os.path.join(HERE, "orb_golden.npz"), **out)
    for k_, v in out.items():
        print(k_, v.shape, v.dtype)


if __name__ == "__main__":
    main()
