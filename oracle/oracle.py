"""ctypes front-end to the CPU oracle (oracle/_build/liborb_oracle.so).

ORACLE — TEST INFRASTRUCTURE ONLY. Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never from mcvslam_b200/ (the product path has no CPU fallback).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liborb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
DM_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
assert KP_DTYPE.itemsize == 28 and DM_DTYPE.itemsize == 16


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("orb_oracle.cpp", "ora_primitives.hpp", "ora_lk.hpp", "ora_pattern.inc", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
        L.ora_orb_create.restype = vp
        L.ora_orb_create.argtypes = [i, f, i, i, i]
        L.ora_orb_destroy.argtypes = [vp]
        L.ora_orb_set_debug.argtypes = [vp, i]
        L.ora_orb_params.argtypes = [vp] + [vp] * 6
        L.ora_orb_extract.argtypes = [vp, vp, i, i, i, vp, i, vp, i]
        L.ora_orb_level_size.argtypes = [vp, i, vp, vp]
        L.ora_orb_level_copy.argtypes = [vp, i, vp]
        L.ora_orb_blurred_copy.argtypes = [vp, i, vp]
        L.ora_orb_debug_kps.argtypes = [vp, i, i, vp, i]
        L.ora_distribute_octree.argtypes = [vp, i, i, i, i, i, i, vp, i]
        L.ora_resize_linear_u8.argtypes = [vp, i, i, i, vp, i, i, i]
        L.ora_gauss7_u8.argtypes = [vp, i, i, i, vp, i]
        L.ora_bgr2gray_u8.argtypes = [vp, i, i, i, vp, i]
        L.ora_fast_atan2.restype = f
        L.ora_fast_atan2.argtypes = [f, f]
        L.ora_fast_atan2_array.argtypes = [vp, vp, vp, i]
        L.ora_sincosf_array.argtypes = [vp, vp, vp, i]
        L.ora_fast.argtypes = [vp, i, i, i, i, vp, i]
        L.ora_hamming256.argtypes = [vp, vp]
        L.ora_knn2_firstparty.argtypes = [vp, i, vp, i, vp]
        L.ora_knn2_bf.argtypes = [vp, i, vp, i, vp]
        L.ora_knn2_candidates.argtypes = [vp, i, vp, vp, vp, vp]
        L.ora_filter_ratio.argtypes = [vp, i, i, f, vp]
        L.ora_filter_threshold.argtypes = [vp, i, i]
        L.ora_filter_orientation.argtypes = [vp, i, vp, vp]
        L.ora_stereo_match.argtypes = [vp, vp, vp, vp, i, vp, vp, i, i, f, f, vp, vp, vp, vp]
        L.ora_project_match.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp, vp, vp, vp, i, f, vp, vp]
        L.ora_bench_frames.restype = d
        L.ora_bench_frames.argtypes = [vp, i, i, i, i, f, i, i, i, f, f, i, i, vp]
        L.ora_bench_knn2.restype = d
        L.ora_bench_knn2.argtypes = [vp, i, vp, i, i, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a


def resize_linear(img, dw, dh):
    img = _u8(img)
    out = np.empty((dh, dw), np.uint8)
    lib().ora_resize_linear_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), dw, dh, dw)
    return out


def bgr2gray(img):
    """cv::cvtColor(COLOR_BGR2GRAY), System::Track (src/System.cpp:60-64)."""
    img = _u8(img)
    h, w, c = img.shape
    assert c == 3
    out = np.empty((h, w), np.uint8)
    lib().ora_bgr2gray_u8(_p(img), w, h, img.strides[0], _p(out), out.strides[0])
    return out


def gauss7(img):
    img = _u8(img)
    out = np.empty_like(img)
    lib().ora_gauss7_u8(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), img.shape[1])
    return out


def fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32); x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(y)
    lib().ora_fast_atan2_array(_p(y), _p(x), _p(out), y.size)
    return out


def sincosf(a):
    a = np.ascontiguousarray(a, np.float32)
    s = np.empty_like(a); c = np.empty_like(a)
    lib().ora_sincosf_array(_p(a), _p(s), _p(c), a.size)
    return s, c


def fast(img, threshold):
    """cv::FAST(img, th, nonmax=True) model: returns int array (n,3) of x, y, score, row-major order."""
    img = _u8(img)
    cap = max(16, img.size // 4)
    out = np.empty((cap, 3), np.int32)
    n = lib().ora_fast(_p(img), img.shape[1], img.shape[0], img.strides[0], threshold, _p(out), cap)
    return out[:n].copy()


class Orb:
    """Mirror of ORB_SLAM3::ORBextractor as used by MCVSLAM::ORB (ORBExtractor.cpp:10-38)."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=28, min_th=15, debug=False):
        self.h = lib().ora_orb_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        if debug:
            lib().ora_orb_set_debug(self.h, 1)
        sc = [np.empty(nlevels, np.float32) for _ in range(4)]
        q = np.empty(nlevels, np.int32); um = np.empty(16, np.int32)
        lib().ora_orb_params(self.h, *[_p(a) for a in sc], _p(q), _p(um))
        self.scale, self.inv_scale, self.sigma2, self.inv_sigma2 = sc
        self.quota, self.umax = q, um

    def __del__(self):
        try:
            lib().ora_orb_destroy(self.h)
        except Exception:
            pass

    def extract(self, img, seeds=None):
        img = _u8(img)
        cap = self.nfeatures + 4 * self.nlevels + 64 + (0 if seeds is None else len(seeds))
        kps = np.zeros(cap, KP_DTYPE)
        ns = 0
        if seeds is not None and len(seeds):
            ns = len(seeds); kps[:ns] = seeds
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().ora_orb_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), ns, _p(desc), cap)
        if n < 0:
            return n, None, None
        return n, kps[:n].copy(), desc[:n].copy()

    def level(self, l):
        w = C.c_int(); h = C.c_int()
        lib().ora_orb_level_size(self.h, l, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        lib().ora_orb_level_copy(self.h, l, _p(out))
        return out

    def blurred(self, l):
        lv = self.level(l)
        out = np.zeros_like(lv)
        lib().ora_orb_blurred_copy(self.h, l, _p(out))
        return out

    def debug_kps(self, which, l):
        cap = 1 << 17
        out = np.zeros(cap, KP_DTYPE)
        n = lib().ora_orb_debug_kps(self.h, which, l, _p(out), cap)
        return out[:n].copy()


def distribute_octree(kps, min_x, max_x, min_y, max_y, n_target):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    cap = len(kps) + 8
    out = np.zeros(cap, KP_DTYPE)
    n = lib().ora_distribute_octree(_p(kps), len(kps), min_x, max_x, min_y, max_y, n_target, _p(out), cap)
    return out[:n].copy()


def knn2_firstparty(q, t):
    q = _u8(q); t = _u8(t)
    out = np.zeros((len(q), 2), DM_DTYPE)
    lib().ora_knn2_firstparty(_p(q), len(q), _p(t), len(t), _p(out))
    return out


def knn2_bf(q, t):
    q = _u8(q); t = _u8(t)
    out = np.zeros((len(q), 2), DM_DTYPE)
    k = lib().ora_knn2_bf(_p(q), len(q), _p(t), len(t), _p(out))
    return out, k


def knn2_candidates(q, t, cand_off, cand_idx):
    q = _u8(q); t = _u8(t)
    cand_off = np.ascontiguousarray(cand_off, np.int32); cand_idx = np.ascontiguousarray(cand_idx, np.int32)
    out = np.zeros((len(q), 2), DM_DTYPE)
    lib().ora_knn2_candidates(_p(q), len(q), _p(t), _p(cand_off), _p(cand_idx), _p(out))
    return out


def filter_ratio(knn, ratio=0.6):
    knn = np.ascontiguousarray(knn, DM_DTYPE)
    nq, per = knn.shape
    out = np.zeros(nq, DM_DTYPE)
    n = lib().ora_filter_ratio(_p(knn), nq, per, ratio, _p(out))
    return out[:n].copy()


def filter_threshold(m, th=46):
    m = np.ascontiguousarray(m, DM_DTYPE).copy()
    n = lib().ora_filter_threshold(_p(m), len(m), th)
    return m[:n].copy()


def filter_orientation(m, kps1, kps2):
    m = np.ascontiguousarray(m, DM_DTYPE).copy()
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    n = lib().ora_filter_orientation(_p(m), len(m), _p(kps1), _p(kps2))
    return m[:n].copy()


def filter_fmatrix(m, kps1, kps2, F12, level_sigma2):
    """MatchRes::FilterFMatrix (src/Matcher.cpp:76-91,310-325): swap-remove by the epipolar distance test."""
    m = np.ascontiguousarray(m, DM_DTYPE).copy()
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    F = np.ascontiguousarray(F12, np.float32).reshape(9); s2 = np.ascontiguousarray(level_sigma2, np.float32)
    L = lib()
    L.ora_filter_fmatrix.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
    n = L.ora_filter_fmatrix(_p(m), len(m), _p(kps1), _p(kps2), _p(F), _p(s2))
    return m[:n].copy()


def stereo_match(orb_l, orb_r, kl, dl, kr, dr, n_rows, bf, b):
    kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE)
    dl = _u8(dl); dr = _u8(dr)
    ur = np.empty(len(kl), np.float32); dp = np.empty(len(kl), np.float32)
    bd = np.empty(len(kl), np.int32); br = np.empty(len(kl), np.int32)
    n = lib().ora_stereo_match(orb_l.h, orb_r.h, _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), n_rows, bf, b,
                               _p(ur), _p(dp), _p(bd), _p(br))
    return n, ur, dp, bd, br


def project_match(kps, desc, W, H, scale_factors, Rcw, tcw, intr, mp_xyz, mp_desc, mp_level, r_threshold):
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = _u8(desc)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    Rcw = np.ascontiguousarray(Rcw, np.float32); tcw = np.ascontiguousarray(tcw, np.float32)
    intr = np.ascontiguousarray(intr, np.float32)
    mp_xyz = np.ascontiguousarray(mp_xyz, np.float32); mp_desc = _u8(mp_desc)
    mp_level = np.ascontiguousarray(mp_level, np.int32)
    n_mp = len(mp_level)
    oi = np.empty(n_mp, np.int32); od = np.empty(n_mp, np.int32)
    cnt = lib().ora_project_match(_p(kps), _p(desc), len(kps), W, H, _p(sf), _p(Rcw), _p(tcw), _p(intr), _p(mp_xyz),
                                  _p(mp_desc), _p(mp_level), n_mp, r_threshold, _p(oi), _p(od))
    return cnt, oi, od


def fuse_match(kps, desc, W, H, sigma2, inv_sigma2, Rcw, tcw, Ow, intr, depth_left, bf, mp_xyz, mp_normal, mp_desc, mp_level):
    """Map::Fuse matching front-end (src/Map.cpp:478-527). Returns (cnt, out_idx, out_dist)."""
    kps = np.ascontiguousarray(kps, KP_DTYPE); desc = _u8(desc)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    sigma2, inv_sigma2, Rcw, tcw, Ow, intr, depth_left, mp_xyz, mp_normal = map(f32, (sigma2, inv_sigma2, Rcw, tcw, Ow, intr, depth_left, mp_xyz, mp_normal))
    mp_desc = _u8(mp_desc); mp_level = np.ascontiguousarray(mp_level, np.int32)
    n_mp = len(mp_level)
    oi = np.empty(n_mp, np.int32); od = np.empty(n_mp, np.int32)
    L = lib()
    L.ora_fuse_match.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 7 + [C.c_float] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p]
    cnt = L.ora_fuse_match(_p(kps), _p(desc), len(kps), W, H, _p(sigma2), _p(inv_sigma2), _p(Rcw), _p(tcw), _p(Ow), _p(intr), _p(depth_left),
                           bf, _p(mp_xyz), _p(mp_normal), _p(mp_desc), _p(mp_level), n_mp, _p(oi), _p(od))
    return cnt, oi, od


def wnd_track(kps1, desc1, q_idx, kps2, desc2, W, H):
    """Tracker::Wnd_Track (src/Tracker.cpp:341-360). Returns (cnt, out_idx [reference behaviour: first candidate], out_best, out_dist)."""
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); desc1 = _u8(desc1); kps2 = np.ascontiguousarray(kps2, KP_DTYPE); desc2 = _u8(desc2)
    q_idx = np.ascontiguousarray(q_idx, np.int32)
    n_q = len(q_idx)
    oi = np.empty(n_q, np.int32); ob = np.empty(n_q, np.int32); od = np.empty(n_q, np.int32)
    L = lib()
    L.ora_wnd_track.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    cnt = L.ora_wnd_track(_p(kps1), _p(desc1), _p(q_idx), n_q, _p(kps2), _p(desc2), len(kps2), W, H, _p(oi), _p(ob), _p(od))
    return cnt, oi, ob, od


def distinctive(desc, off):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cpp:101-150) for a batch: rows [off[m], off[m+1]) of desc are the
    observations of point m. Returns (best_idx, best_median), -1 for points without observations."""
    desc = _u8(desc); off = np.ascontiguousarray(off, np.int32)
    n_mp = len(off) - 1
    bi = np.empty(n_mp, np.int32); bm = np.empty(n_mp, np.int32)
    L = lib()
    L.ora_distinctive.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ora_distinctive.restype = None
    L.ora_distinctive(_p(desc), _p(off), n_mp, _p(bi), _p(bm))
    return bi, bm


def lk_track(prev, nxt, pts):
    """cv::calcOpticalFlowPyrLK as KL_Track calls it (src/Frame.cpp:52-54). pts [n, 2] float32. Returns (next_pts, status, err)."""
    prev = np.ascontiguousarray(prev, np.uint8); nxt = np.ascontiguousarray(nxt, np.uint8)
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2); n = len(pts)
    out = np.zeros((n, 2), np.float32); st = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
    L = lib()
    L.ora_lk_track.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ora_lk_track.restype = None
    L.ora_lk_track(_p(prev), _p(nxt), prev.shape[1], prev.shape[0], prev.strides[0], _p(pts), n, _p(out), _p(st), _p(err))
    return out, st, err


def kl_track(prev, nxt, kps):
    """KL_Track (src/Frame.cpp:34-76) without the MapPoint map: returns (cnt, new_kps, ok, next_pts, status, err)."""
    prev = np.ascontiguousarray(prev, np.uint8); nxt = np.ascontiguousarray(nxt, np.uint8)
    kps = np.ascontiguousarray(kps, KP_DTYPE); n = len(kps)
    new = np.zeros(n, KP_DTYPE); ok = np.zeros(n, np.uint8)
    out = np.zeros((n, 2), np.float32); st = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
    L = lib()
    L.ora_kl_track.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    cnt = L.ora_kl_track(_p(prev), _p(nxt), prev.shape[1], prev.shape[0], prev.strides[0], _p(kps), n, _p(new), _p(ok), _p(out), _p(st), _p(err))
    return cnt, new, ok, out, st, err


def bow_transform(desc, voc, levelsup=4):
    """DBoW3::Vocabulary::transform (modules/DBow3/src/Vocabulary.cpp:572-672) as Object::ComputeBow calls it. `voc` = dict of the
    flat vocabulary arrays (mcvslam_b200.synth.random_vocabulary layout). Returns dict(word, weight, nid, bow_ids, bow_vals,
    fv_nodes, fv_off, fv_idx)."""
    desc = _u8(desc); n = len(desc)
    co = np.ascontiguousarray(voc["child_off"], np.int32); ci = np.ascontiguousarray(voc["child_ids"], np.uint32)
    nd = _u8(voc["node_desc"]); wi = np.ascontiguousarray(voc["word_id"], np.int32); ww = np.ascontiguousarray(voc["weight"], np.float64)
    ow = np.zeros(n, np.int32); owt = np.zeros(n, np.float64); onid = np.zeros(n, np.uint32)
    bi = np.zeros(n + 1, np.uint32); bv = np.zeros(n + 1, np.float64); fn = np.zeros(n + 1, np.uint32); fo = np.zeros(n + 2, np.int32); fi = np.zeros(n + 1, np.int32)
    nfv = C.c_int(0)
    L = lib()
    L.ora_bow_transform.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p] * 8 + [C.POINTER(C.c_int)]
    k = L.ora_bow_transform(_p(desc), n, _p(co), _p(ci), _p(nd), _p(wi), _p(ww), int(voc["L"]), levelsup, int(voc["weighting"]), int(voc["norm"]),
                            _p(ow), _p(owt), _p(onid), _p(bi), _p(bv), _p(fn), _p(fo), _p(fi), C.byref(nfv))
    m = nfv.value
    return dict(word=ow, weight=owt, nid=onid, bow_ids=bi[:k].copy(), bow_vals=bv[:k].copy(), fv_nodes=fn[:m].copy(), fv_off=fo[:m + 1].copy(),
                fv_idx=fi[:fo[m]].copy())


def bench_frames(imgs, nfeatures, sf, nlevels, ini, mn, bf, b, n_threads, repeat=1):
    """imgs: (n_frames, 3, H, W) u8. Returns (seconds, total_keypoints)."""
    imgs = _u8(imgs)
    n, three, h, w = imgs.shape
    assert three == 3
    tk = C.c_longlong(0)
    s = lib().ora_bench_frames(_p(imgs), n, w, h, nfeatures, sf, nlevels, ini, mn, bf, b, n_threads, repeat, C.byref(tk))
    return s, tk.value


def bench_knn2(q, t, n_threads):
    q = _u8(q); t = _u8(t)
    out = np.zeros((len(q), 2), DM_DTYPE)
    s = lib().ora_bench_knn2(_p(q), len(q), _p(t), len(t), n_threads, _p(out))
    return s, out
