"""ctypes front-end to oracle/_ref/libmcv_ref.so — the REFERENCE'S OWN sources (ORBextractor.cc, Matcher.cpp, Frame.cpp,
Object.cpp, MapPoint.cpp, Map.cpp, Tracker.cpp, DBoW3 ...) compiled unmodified by oracle/build_ref.py against oracle/ref_shim.

ORACLE — TEST INFRASTRUCTURE ONLY. Used by tests/ (oracle == _ref pins the restatement; CUDA == _ref on the GPU box) and by
bench.py's cpu_baseline / --impl reference legs (`kind: "reference"`); never by mcvslam_b200/.
"""
import ctypes as C
import os
import tempfile

import numpy as np

from . import build_ref
from .oracle import DM_DTYPE, KP_DTYPE

_lib = None
_tmp = None


def available():
    """True when the .so exists or can be built (reference sources present)."""
    return os.path.exists(build_ref.SO) or build_ref.available()


def build(force=False):
    return build_ref.build(force)


def lib():
    global _lib
    if _lib is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref/libmcv_ref.so is missing and /root/reference is not present to build it")
        L = C.CDLL(so)
        vp, i, f, d = C.c_void_p, C.c_int, C.c_float, C.c_double
        L.ref_version.restype = C.c_char_p
        L.ref_orb_create.restype = vp
        L.ref_orb_create.argtypes = [C.c_char_p]
        L.ref_orb_destroy.argtypes = [vp]
        L.ref_orb_levels.argtypes = [vp]
        L.ref_orb_params.argtypes = [vp] * 7
        L.ref_orb_extract.argtypes = [vp, vp, i, i, i, vp, i, vp, i]
        L.ref_orb_level_size.argtypes = [vp, i, vp, vp]
        L.ref_orb_level_copy.argtypes = [vp, i, vp]
        L.ref_orb_level_bordered_copy.argtypes = [vp, i, vp]
        L.ref_distribute_octree.argtypes = [vp, i, i, i, i, i, i, vp, i]
        L.ref_hamming.restype = C.c_uint
        L.ref_hamming.argtypes = [vp, vp]
        L.ref_knn2_firstparty.argtypes = [vp, i, vp, i, vp]
        L.ref_knn2_bf.argtypes = [vp, i, vp, i, vp, i]
        L.ref_bf_match.argtypes = [vp, i, vp, i, vp]
        L.ref_knn2_candidates.argtypes = [vp, i, vp, vp, vp, vp]
        L.ref_filter_ratio.argtypes = [vp, i, i, f, vp]
        L.ref_filter_threshold.argtypes = [vp, i, i]
        L.ref_filter_orientation.argtypes = [vp, i, vp, i, vp, i]
        L.ref_filter_fmatrix.argtypes = [vp, i, vp, i, vp, i, vp, vp, i]
        L.ref_dbow_match.argtypes = [vp, i, vp, vp, vp, i, vp, i, vp, vp, vp, i, vp, i]
        L.ref_rig_create.restype = vp
        L.ref_rig_create.argtypes = [C.c_char_p, f, f]
        L.ref_rig_destroy.argtypes = [vp]
        L.ref_rig_frame.argtypes = [vp, vp, i, i, vp, vp, vp, vp, vp, i]
        L.ref_rig_extract_lr.argtypes = [vp, vp, vp, i, i]
        L.ref_rig_stereo.argtypes = [vp, vp, vp, i, vp, vp, i, i, i, vp, vp]
        L.ref_rig_bench.restype = d
        L.ref_rig_bench.argtypes = [vp, vp, i, i, i, i, vp, vp]
        L.ref_obj_create.restype = vp
        L.ref_obj_create.argtypes = [vp, vp, vp, i, i, i, vp, vp, vp]
        L.ref_obj_destroy.argtypes = [vp]
        L.ref_obj_features_in_area.argtypes = [vp, f, f, f, vp, i]
        L.ref_project_match.argtypes = [vp, vp, vp, vp, i, f, vp]
        L.ref_voc_load.argtypes = [C.c_char_p]
        L.ref_voc_save.argtypes = [C.c_char_p]
        L.ref_obj_compute_bow.argtypes = [vp] * 7
        L.ref_distinctive.restype = None
        L.ref_distinctive.argtypes = [vp, vp, vp, i, vp, vp]
        L.ref_distinctive_order.restype = None
        L.ref_distinctive_order.argtypes = [i, vp]
        L.ref_kl_track.argtypes = [vp, vp, vp, i, i, i, vp, i, vp, vp]
        L.ref_fuse_match.argtypes = [vp, C.c_char_p, vp, i, f, vp, vp, vp, vp, i, vp]
        L.ref_compute_f12.restype = None
        L.ref_compute_f12.argtypes = [vp, vp, vp]
        L.ref_wnd_track.argtypes = [vp, vp, C.c_char_p, vp, i, vp]
        L.ref_set_num_threads.argtypes = [i]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _f32(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def _tmpdir():
    global _tmp
    if _tmp is None:
        _tmp = tempfile.TemporaryDirectory(prefix="mcv_ref_")
    return _tmp.name


def extractor_yaml(nfeatures, scale_factor, nlevels, ini_th, min_th):
    """Writes an extractor config in the reference's format (config/extractor.yaml) and returns its path."""
    path = os.path.join(_tmpdir(), "extractor_%d_%r_%d_%d_%d.yaml" % (nfeatures, scale_factor, nlevels, ini_th, min_th))
    with open(path, "w") as fh:
        fh.write("nkeypoints: %d\nscale_factor: %r\nnlevels: %d\nORBextractor.iniThFAST: %d\nORBextractor.minThFAST: %d\n" %
                 (nfeatures, scale_factor, nlevels, ini_th, min_th))
    return path


def system_yaml():
    """config/system.yaml keys that Map / Tracker constructors read (src/Map.cpp:25-30, src/Tracker.cpp:24-32)."""
    path = os.path.join(_tmpdir(), "system.yaml")
    with open(path, "w") as fh:
        fh.write("connection_threshold: 30\nmappoint_life_span: 7\nTh_depth: 10\nTh_motionmodel_min_mps: 40\nTh_local_map_min_mps: 50\n"
                 "Th_lastkeyframe_min_mps: 50\nTh_max_frame_interval: 15\n")
    return path


class Orb:
    """MCVSLAM::ORB(config_path) of the reference (ORBExtractor.cpp:20-23)."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=28, min_th=15):
        self.h = lib().ref_orb_create(extractor_yaml(nfeatures, scale_factor, nlevels, ini_th, min_th).encode())
        assert self.h, "reference ORB construction failed"
        self.nlevels, self.nfeatures = nlevels, nfeatures
        sc = [np.empty(nlevels, np.float32) for _ in range(4)]
        q = np.empty(nlevels, np.int32); um = np.empty(16, np.int32)
        lib().ref_orb_params(self.h, *[_p(a) for a in sc], _p(q), _p(um))
        self.scale, self.inv_scale, self.sigma2, self.inv_sigma2 = sc
        self.quota, self.umax = q, um

    def __del__(self):
        try:
            lib().ref_orb_destroy(self.h)
        except Exception:
            pass

    def extract(self, img, seeds=None):
        img = _u8(img)
        cap = self.nfeatures + 16 * self.nlevels + 64 + (0 if seeds is None else len(seeds))
        kps = np.zeros(cap, KP_DTYPE)
        ns = 0
        if seeds is not None and len(seeds):
            ns = len(seeds); kps[:ns] = seeds
        desc = np.zeros((cap, 32), np.uint8)
        if img.size == 0:
            n = lib().ref_orb_extract(self.h, None, 0, 0, 0, _p(kps), ns, _p(desc), cap)
        else:
            n = lib().ref_orb_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), ns, _p(desc), cap)
        if n < 0:
            return n, None, None
        return n, kps[:n].copy(), desc[:n].copy()

    def level(self, l, bordered=False):
        w = C.c_int(); h = C.c_int()
        lib().ref_orb_level_size(self.h, l, C.byref(w), C.byref(h))
        if bordered:
            out = np.empty((h.value + 38, w.value + 38), np.uint8)
            lib().ref_orb_level_bordered_copy(self.h, l, _p(out))
            return out
        out = np.empty((h.value, w.value), np.uint8)
        lib().ref_orb_level_copy(self.h, l, _p(out))
        return out


def distribute_octree(kps, min_x, max_x, min_y, max_y, n_target):
    kps = np.ascontiguousarray(kps, KP_DTYPE)
    cap = len(kps) + 8
    out = np.zeros(cap, KP_DTYPE)
    n = lib().ref_distribute_octree(_p(kps), len(kps), min_x, max_x, min_y, max_y, n_target, _p(out), cap)
    return out[:n].copy()


def hamming(a, b):
    return int(lib().ref_hamming(_p(_u8(a)), _p(_u8(b))))


def knn2_firstparty(q, t):
    q = _u8(q); t = _u8(t)
    out = np.zeros((len(q), 2), DM_DTYPE)
    lib().ref_knn2_firstparty(_p(q), len(q), _p(t), len(t), _p(out))
    return out


def knn2_bf(q, t, cv_variant=False):
    q = _u8(q); t = _u8(t)
    out = np.zeros((len(q), 2), DM_DTYPE)
    k = lib().ref_knn2_bf(_p(q), len(q), _p(t), len(t), _p(out), 1 if cv_variant else 0)
    return out, k


def bf_match(q, t):
    q = _u8(q); t = _u8(t)
    out = np.zeros(len(q), DM_DTYPE)
    n = lib().ref_bf_match(_p(q), len(q), _p(t), len(t), _p(out))
    return out[:n].copy()


def knn2_candidates(q, t, cand_off, cand_idx):
    q = _u8(q); t = _u8(t)
    cand_off = np.ascontiguousarray(cand_off, np.int32); cand_idx = np.ascontiguousarray(cand_idx, np.int32)
    out = np.zeros((len(q), 2), DM_DTYPE)
    lib().ref_knn2_candidates(_p(q), len(q), _p(t), _p(cand_off), _p(cand_idx), _p(out))
    return out


def filter_ratio(knn, ratio=0.6):
    knn = np.ascontiguousarray(knn, DM_DTYPE)
    nq, per = knn.shape
    out = np.zeros(nq, DM_DTYPE)
    n = lib().ref_filter_ratio(_p(knn), nq, per, ratio, _p(out))
    return out[:n].copy()


def filter_threshold(m, th=46):
    m = np.ascontiguousarray(m, DM_DTYPE).copy()
    n = lib().ref_filter_threshold(_p(m), len(m), th)
    return m[:n].copy()


def filter_orientation(m, kps1, kps2):
    m = np.ascontiguousarray(m, DM_DTYPE).copy()
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    n = lib().ref_filter_orientation(_p(m), len(m), _p(kps1), len(kps1), _p(kps2), len(kps2))
    return m[:n].copy()


def filter_fmatrix(m, kps1, kps2, F12, level_sigma2):
    m = np.ascontiguousarray(m, DM_DTYPE).copy()
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
    F = _f32(F12).reshape(9); s2 = _f32(level_sigma2)
    n = lib().ref_filter_fmatrix(_p(m), len(m), _p(kps1), len(kps1), _p(kps2), len(kps2), _p(F), _p(s2), len(s2))
    return m[:n].copy()


def flat_fv(fv):
    """dict node -> [feature idx] (or the flattened triple) -> (nodes, off, idx) arrays, node ids ascending."""
    if isinstance(fv, tuple):
        return tuple(np.ascontiguousarray(a, dt) for a, dt in zip(fv, (np.uint32, np.int32, np.int32)))
    nodes = sorted(fv)
    off = np.zeros(len(nodes) + 1, np.int32)
    idx = []
    for k, nd in enumerate(nodes):
        idx += list(fv[nd]); off[k + 1] = len(idx)
    return np.array(nodes, np.uint32), off, np.array(idx, np.int32)


def dbow_match(d1, fv1, d2, fv2):
    d1 = _u8(d1); d2 = _u8(d2)
    n1, o1, i1 = flat_fv(fv1); n2, o2, i2 = flat_fv(fv2)
    out = np.zeros((len(d1), 2), DM_DTYPE)
    n = lib().ref_dbow_match(_p(d1), len(d1), _p(n1), _p(o1), _p(i1), len(n1), _p(d2), len(d2), _p(n2), _p(o2), _p(i2), len(n2), _p(out), len(d1))
    return out[:n].copy()


class Rig:
    """The reference's Frame pipeline: static extractors + the real Frame constructor (src/Frame.cpp:24-32,78-138)."""

    def __init__(self, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=28, min_th=15, bf=955.40503, baseline=1.0):
        self.h = lib().ref_rig_create(extractor_yaml(nfeatures, scale_factor, nlevels, ini_th, min_th).encode(), bf, baseline)
        assert self.h
        self.cap = nfeatures + 16 * nlevels + 64

    def __del__(self):
        try:
            lib().ref_rig_destroy(self.h)
        except Exception:
            pass

    def frame(self, triplet):
        """triplet (3, H, W) u8 -> dict(counts, kps [3][cap], desc, u_right, depth_left)."""
        t = _u8(triplet)
        _, h, w = t.shape
        cap = self.cap
        kps = np.zeros((3, cap), KP_DTYPE); desc = np.zeros((3, cap, 32), np.uint8); cnt = np.zeros(3, np.int32)
        ur = np.full(cap, -1, np.float32); dp = np.full(cap, -1, np.float32)
        rc = lib().ref_rig_frame(self.h, _p(t), w, h, _p(kps), _p(desc), _p(cnt), _p(ur), _p(dp), cap)
        assert rc == 0, rc
        return dict(counts=cnt, kps=kps, desc=desc, u_right=ur, depth_left=dp)

    def extract_lr(self, left, right):
        left = _u8(left); right = _u8(right)
        lib().ref_rig_extract_lr(self.h, _p(left), _p(right), left.shape[1], left.shape[0])

    def stereo(self, kl, dl, kr, dr, w, h):
        """Frame::ComputeStereoMatch on caller keypoints against the pyramids of the last frame() / extract_lr()."""
        kl = np.ascontiguousarray(kl, KP_DTYPE); kr = np.ascontiguousarray(kr, KP_DTYPE); dl = _u8(dl); dr = _u8(dr)
        ur = np.empty(len(kl), np.float32); dp = np.empty(len(kl), np.float32)
        lib().ref_rig_stereo(self.h, _p(kl), _p(dl), len(kl), _p(kr), _p(dr), len(kr), w, h, _p(ur), _p(dp))
        return ur, dp

    def bench(self, frames, repeat=1):
        """frames (n, 3, H, W) u8. Returns (seconds, keypoints, dict of the reference's own MyTimer stage seconds)."""
        f = _u8(frames)
        n, _, h, w = f.shape
        st = np.zeros(3, np.float64); kp = C.c_longlong(0)
        s = lib().ref_rig_bench(self.h, _p(f), n, w, h, repeat, _p(st), C.byref(kp))
        return s, kp.value, dict(KL=float(st[0]), ORBE=float(st[1]), SMatch=float(st[2]))


class Obj:
    """MCVSLAM::Object with caller keypoints / descriptors (pose + pinhole intrinsics optional)."""

    def __init__(self, orb, kps, desc, w, h, intr=None, Rcw=None, tcw=None):
        self.orb = orb
        kps = np.ascontiguousarray(kps, KP_DTYPE); desc = _u8(desc)
        self.n = len(kps)
        self.h = lib().ref_obj_create(orb.h, _p(kps), _p(desc), len(kps), w, h, _p(_f32(intr)), _p(_f32(Rcw)), _p(_f32(tcw)))

    def __del__(self):
        try:
            lib().ref_obj_destroy(self.h)
        except Exception:
            pass

    def features_in_area(self, x, y, r):
        out = np.zeros(self.n + 1, np.int32)
        n = lib().ref_obj_features_in_area(self.h, x, y, r, _p(out), len(out))
        return out[:n].copy()

    def project_match(self, mp_xyz, mp_desc, mp_level, r_threshold):
        xyz = _f32(mp_xyz); md = _u8(mp_desc); lv = np.ascontiguousarray(mp_level, np.int32)
        oi = np.empty(len(lv), np.int32)
        cnt = lib().ref_project_match(self.h, _p(xyz), _p(md), _p(lv), len(lv), r_threshold, _p(oi))
        return cnt, oi

    def compute_bow(self):
        n = self.n
        bi = np.zeros(n + 1, np.uint32); bv = np.zeros(n + 1, np.float64); fn = np.zeros(n + 1, np.uint32); fo = np.zeros(n + 2, np.int32)
        fi = np.zeros(n + 1, np.int32); nfv = C.c_int(0)
        k = lib().ref_obj_compute_bow(self.h, _p(bi), _p(bv), _p(fn), _p(fo), _p(fi), C.byref(nfv))
        m = nfv.value
        return dict(bow_ids=bi[:k].copy(), bow_vals=bv[:k].copy(), fv_nodes=fn[:m].copy(), fv_off=fo[:m + 1].copy(), fv_idx=fi[:fo[m]].copy())

    def fuse_match(self, depth_left, bf, mp_xyz, mp_normal, mp_desc, mp_level):
        dl = _f32(depth_left); xyz = _f32(mp_xyz); nr = _f32(mp_normal); md = _u8(mp_desc); lv = np.ascontiguousarray(mp_level, np.int32)
        oi = np.empty(len(lv), np.int32)
        cnt = lib().ref_fuse_match(self.h, system_yaml().encode(), _p(dl), len(dl), bf, _p(xyz), _p(nr), _p(md), _p(lv), len(lv), _p(oi))
        return cnt, oi


def compute_f12(obj1, obj2):
    F = np.zeros((3, 3), np.float32)
    lib().ref_compute_f12(obj1.h, obj2.h, _p(F))
    return F


def wnd_track(obj1, obj2, q_idx):
    q = np.ascontiguousarray(q_idx, np.int32)
    oi = np.empty(len(q), np.int32)
    cnt = lib().ref_wnd_track(obj1.h, obj2.h, system_yaml().encode(), _p(q), len(q), _p(oi))
    return cnt, oi


def write_dbow3_binary(voc, path):
    """Serialises a flat vocabulary (mcvslam_b200.synth.random_vocabulary layout) in DBoW3's own uncompressed binary stream
    format (modules/DBow3/src/Vocabulary.cpp:1076-1141 fromStream) so that the reference's loader reads it."""
    import struct
    co = np.asarray(voc["child_off"]); ci = np.asarray(voc["child_ids"]); nd = _u8(voc["node_desc"])
    wi = np.asarray(voc["word_id"]); ww = np.asarray(voc["weight"], np.float64)
    n = len(wi)
    parent = np.zeros(n, np.uint32)
    for p in range(n):
        for c in ci[co[p]:co[p + 1]]:
            parent[c] = p
    scoring = {0: 5, 1: 0, 2: 1}[int(voc["norm"])] if "scoring" not in voc else int(voc["scoring"])   # DOT_PRODUCT (no norm), L1_NORM, L2_NORM
    k = int(max(co[1:] - co[:-1]))
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", 88877711233)); f.write(struct.pack("<?", False)); f.write(struct.pack("<I", n))
        f.write(struct.pack("<iiii", k, int(voc["L"]), scoring, int(voc["weighting"])))
        # nodes must come in an order where children follow parents and siblings keep their stored order
        from collections import deque
        order = []
        dq = deque([0])
        while dq:
            p = dq.popleft()
            for c in ci[co[p]:co[p + 1]]:
                order.append(int(c)); dq.append(int(c))
        # fromStream pushes each node onto its parent's children in FILE order: emit parent by parent, siblings in stored order
        for c in order:
            f.write(struct.pack("<I", c)); f.write(struct.pack("<I", int(parent[c]))); f.write(struct.pack("<d", float(ww[c])))
            f.write(struct.pack("<iii", 32, 1, 0)); f.write(nd[c].tobytes())       # DescManip::toStream: cols, rows, type (CV_8UC1 = 0), data
        words = [(int(wi[i]), i) for i in range(n) if wi[i] >= 0]
        words.sort()
        f.write(struct.pack("<I", len(words)))
        for w, nid in words:
            f.write(struct.pack("<II", w, nid))
    return path


def voc_load(path):
    return lib().ref_voc_load(path.encode())


def voc_save_uncompressed(path):
    """Vocabulary::save(path, false) of the vocabulary the reference currently holds (DBoW3's uncompressed binary stream)."""
    return lib().ref_voc_save(path.encode())


def read_dbow3_binary(path):
    """Inverse of write_dbow3_binary: DBoW3's UNCOMPRESSED binary stream (Vocabulary.cpp:935-1000 toStream) -> the flat arrays
    mcv_voc_create / oracle.bow_transform take. Children keep the order fromStream gives them (file order)."""
    raw = np.fromfile(path, np.uint8)
    import struct
    sig, = struct.unpack_from("<Q", raw, 0)
    assert sig == 88877711233 and raw[8] == 0, "not an uncompressed DBoW3 binary vocabulary"
    n, = struct.unpack_from("<I", raw, 9)
    k, L, scoring, weighting = struct.unpack_from("<iiii", raw, 13)
    rec = np.dtype([("nid", "<u4"), ("parent", "<u4"), ("weight", "<f8"), ("cols", "<i4"), ("rows", "<i4"), ("type", "<i4"), ("desc", "u1", 32)])
    nodes = np.frombuffer(raw, rec, n - 1, 29)
    assert (nodes["cols"] == 32).all() and (nodes["rows"] == 1).all() and (nodes["type"] == 0).all()
    off = 29 + (n - 1) * rec.itemsize
    nw, = struct.unpack_from("<I", raw, off)
    words = np.frombuffer(raw, np.dtype([("wid", "<u4"), ("nid", "<u4")]), nw, off + 4)
    node_desc = np.zeros((n, 32), np.uint8); node_desc[nodes["nid"]] = nodes["desc"]
    weight = np.zeros(n, np.float64); weight[nodes["nid"]] = nodes["weight"]
    word_id = np.full(n, -1, np.int32); word_id[words["nid"]] = words["wid"].astype(np.int32)
    # children of every node in file order: a stable sort by parent keeps it
    order = np.argsort(nodes["parent"], kind="stable")
    child_ids = nodes["nid"][order].astype(np.uint32)
    child_off = np.zeros(n + 1, np.int32)
    np.cumsum(np.bincount(nodes["parent"], minlength=n), out=child_off[1:])
    norm = {0: 1, 1: 2, 2: 1, 3: 1, 4: 1, 5: 0}[int(scoring)]   # ScoringObject.h:73-88: L1, L2, CHI_SQUARE / KL / BHATTACHARYYA (L1), DOT_PRODUCT (none)
    return dict(child_off=child_off, child_ids=child_ids, node_desc=node_desc, word_id=word_id, weight=weight, L=int(L), K=int(k),
                weighting=int(weighting), norm=norm, scoring=int(scoring))


def distinctive(orb, desc, off):
    """MapPoint::ComputeDistinctiveDescriptors per point. Returns (best_row [row within the point's range, -1 none], out_desc)."""
    desc = _u8(desc); off = np.ascontiguousarray(off, np.int32)
    n_mp = len(off) - 1
    br = np.empty(n_mp, np.int32); od = np.zeros((n_mp, 32), np.uint8)
    lib().ref_distinctive(orb.h, _p(desc), _p(off), n_mp, _p(br), _p(od))
    return br, od


def kl_track(orb, prev, nxt, kps):
    """KL_Track (src/Frame.cpp:34-76). Returns (cnt, new_kps [cnt], src [cnt] = input keypoint each new keypoint came from)."""
    prev = _u8(prev); nxt = _u8(nxt); kps = np.ascontiguousarray(kps, KP_DTYPE)
    new = np.zeros(len(kps) + 1, KP_DTYPE); src = np.full(len(kps) + 1, -1, np.int32)
    cnt = lib().ref_kl_track(orb.h, _p(prev), _p(nxt), prev.shape[1], prev.shape[0], prev.strides[0], _p(kps), len(kps), _p(new), _p(src))
    return cnt, new[:cnt].copy(), src[:cnt].copy()


# ---- CPU baseline legs of bench.py (the reference timed on the box's host cores) ------------------------------------------
def bench_frames(cfg, n_procs, n_distinct, seed0, w, h, repeat, cv_threads=1, timeout_s=600):
    """n_procs concurrent replicas of the unmodified reference, one PROCESS each (the three extractors are process-wide statics,
    src/Frame.cpp:24-26, so replicas cannot share a process); every replica constructs n_distinct * repeat Frames from the same
    seeded synthetic triplets, and every Frame fans its three extractions out over the reference's own ThreadPool(3).
    Returns (frames/s over the union of the replicas' timed windows, total frames, stage seconds of replica 0)."""
    import json
    import subprocess
    import sys
    import time
    start_at = time.time() + (0.0 if n_procs <= 1 else 4.0 + 0.1 * n_procs)
    job = json.dumps(dict(cfg=list(cfg), n_distinct=n_distinct, seed0=seed0, w=w, h=h, repeat=repeat, start_at=start_at, cv_threads=cv_threads))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = [subprocess.Popen([sys.executable, "-m", "oracle.ref", job], cwd=root, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
             for _ in range(max(1, n_procs))]
    res = []
    for p in procs:
        out, _ = p.communicate(timeout=timeout_s)
        res.append(json.loads(out.strip().splitlines()[-1]))
    t0 = min(r["t0"] for r in res); t1 = max(r["t1"] for r in res)
    n = sum(r["frames"] for r in res)
    return n / (t1 - t0), n, res[0]["stages"]


def _worker_main(job):
    import json
    import sys
    import time
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mcvslam_b200 import synth
    j = json.loads(job)
    lib().ref_set_num_threads(int(j["cv_threads"]))
    frames = np.stack([synth.triplet(j["seed0"] + s, j["w"], j["h"]) for s in range(j["n_distinct"])])
    rig = Rig(*j["cfg"])
    rig.bench(frames[:1], 1)                       # warm-up: page in, allocate the pyramids
    while time.time() < j["start_at"]:
        time.sleep(0.001)
    t0 = time.time()
    s, kp, stages = rig.bench(frames, int(j["repeat"]))
    print(json.dumps(dict(t0=t0, t1=time.time(), frames=len(frames) * int(j["repeat"]), kp=kp, stages=stages)), flush=True)


if __name__ == "__main__":
    import sys
    _worker_main(sys.argv[1])
