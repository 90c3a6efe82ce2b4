"""ctypes binding of libmcv_b200.so (include/mcv_b200.h) with the reference's class / method names.

This is what tests/ and bench.py drive; the C++ mirror of the same surface is mcvslam_b200/host/mcvslam_b200.hpp.
Nothing here computes: every method forwards host (or device) buffers to the C ABI, which launches the sm_100a kernels.
If the shared library is missing this module raises — there is no CPU fallback (the oracle under oracle/ is test
infrastructure and is never imported from here).

    ORB            <-> MCVSLAM::ORB / ORB_SLAM3::ORBextractor      (ORBExtractor.hpp:8-18, ORBextractor.h:44-99)
    Matcher        <-> MCVSLAM::Matcher statics                    (include/Matcher.hpp:58-92)
    MatchRes(Knn)  <-> MatchRes / MatchResKnn filter chains        (include/Matcher.hpp:39-56)
    Rig            <-> Frame ctor stages ORBE + SMatch             (src/Frame.cpp:118-137)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libmcv_b200.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
DM_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
ORB_GOOD_THRESHOLD = 46  # include/Matcher.hpp:14

EXPORTS = [
    "mcv_last_error", "mcv_version", "mcv_device_count", "mcv_orb_create", "mcv_orb_destroy", "mcv_orb_get_scales",
    "mcv_orb_max_keypoints", "mcv_orb_max_keypoints_for", "mcv_rig_max_keypoints_for", "mcv_orb_extract", "mcv_orb_extract_batch", "mcv_orb_extract_batch_async", "mcv_orb_download_level", "mcv_orb_level_device",
    "mcv_orb_distribute_octree", "mcv_knn2_bf", "mcv_bf_match", "mcv_knn2_firstparty", "mcv_knn2_candidates",
    "mcv_filter_ratio", "mcv_filter_threshold", "mcv_filter_orientation", "mcv_filter_fmatrix", "mcv_dbow_match",
    "mcv_knn2_bf_device", "mcv_knn2_pairs_device", "mcv_rig_create", "mcv_rig_destroy", "mcv_rig_max_keypoints", "mcv_rig_extractor",
    "mcv_rig_set_chunk_frames", "mcv_rig_process", "mcv_rig_set_input_channels", "mcv_orb_set_input_channels", "mcv_rig_process_async", "mcv_rig_submit", "mcv_rig_wait", "mcv_rig_join", "mcv_rig_sync", "mcv_rig_last_launches", "mcv_rig_set_profiling",
    "mcv_rig_stage_ms", "mcv_stereo_match",
    "mcv_project_match", "mcv_fuse_match", "mcv_wnd_track", "mcv_distinctive_descriptors", "mcv_lk_track", "mcv_lk_track_batch", "mcv_kl_track", "mcv_voc_create", "mcv_voc_destroy", "mcv_bow_transform", "mcv_debug_sincosf", "mcv_debug_fast_atan2", "mcv_debug_level_keypoints",
    "mcv_debug_download_blurred", "mcv_debug_popc_peak", "mcv_debug_octree_clocks",
]


class McvError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("mcv_b200 status %d: %s" % (status, msg))
        self.status = status


class OrbParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32)]


class RigParams(C.Structure):
    _fields_ = [("orb", OrbParams), ("bf", C.c_float), ("baseline", C.c_float)]


_lib = None


def lib():
    """Loads the engine. Raises if it has not been built — the product has no other compute path."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("libmcv_b200.so is not built (run `python -m mcvslam_b200.build`); there is no CPU fallback")
        L = C.CDLL(SO_PATH)
        vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        L.mcv_last_error.restype = C.c_char_p
        L.mcv_version.restype = C.c_char_p
        L.mcv_orb_create.argtypes = [C.POINTER(OrbParams), i, vp, C.POINTER(vp)]
        L.mcv_orb_destroy.argtypes = [vp]
        L.mcv_orb_destroy.restype = None
        L.mcv_orb_get_scales.argtypes = [vp] * 6
        L.mcv_orb_max_keypoints.argtypes = [vp, i]
        L.mcv_orb_extract_batch_async.argtypes = [vp, vp, i, i, i, vp, vp, vp, i]
        L.mcv_orb_max_keypoints_for.argtypes = [vp, i, i, i]
        L.mcv_rig_max_keypoints_for.argtypes = [vp, i, i]
        L.mcv_orb_extract.argtypes = [vp, vp, i, i, sz, vp, i, vp, vp, i, C.POINTER(i)]
        L.mcv_orb_extract_batch.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, i, i]
        L.mcv_orb_download_level.argtypes = [vp, i, i, vp, sz, C.POINTER(i), C.POINTER(i)]
        L.mcv_orb_distribute_octree.argtypes = [vp, vp, i, i, i, i, i, i, vp, i, C.POINTER(i)]
        L.mcv_knn2_bf.argtypes = [vp, i, vp, i, vp, C.POINTER(i)]
        L.mcv_bf_match.argtypes = [vp, i, vp, i, vp]
        L.mcv_knn2_firstparty.argtypes = [vp, i, vp, i, vp]
        L.mcv_knn2_candidates.argtypes = [vp, i, vp, i, vp, vp, vp]
        L.mcv_filter_ratio.argtypes = [vp, i, i, f, vp, C.POINTER(i)]
        L.mcv_filter_threshold.argtypes = [vp, C.POINTER(i), i]
        L.mcv_filter_orientation.argtypes = [vp, C.POINTER(i), vp, i, vp, i]
        L.mcv_filter_fmatrix.argtypes = [vp, C.POINTER(i), vp, i, vp, i, vp, vp, i]
        L.mcv_dbow_match.argtypes = [vp, i, vp, vp, vp, i, vp, i, vp, vp, vp, i, vp, C.POINTER(i)]
        L.mcv_knn2_bf_device.argtypes = [vp, i, vp, i, i, vp, vp, vp]
        L.mcv_knn2_pairs_device.argtypes = [vp, vp, i, i, vp, vp, i, vp, vp, vp]
        L.mcv_rig_create.argtypes = [C.POINTER(RigParams), i, vp, C.POINTER(vp)]
        L.mcv_rig_destroy.argtypes = [vp]
        L.mcv_rig_destroy.restype = None
        L.mcv_rig_max_keypoints.argtypes = [vp]
        L.mcv_rig_extractor.argtypes = [vp]
        L.mcv_rig_extractor.restype = vp
        L.mcv_rig_process.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp, vp, i, i]
        L.mcv_rig_process_async.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp, vp, i]
        L.mcv_rig_submit.argtypes = [vp, vp, i, i, i, vp, vp, vp, vp, vp, i, C.POINTER(C.c_longlong)]
        L.mcv_rig_wait.argtypes = [vp, C.c_longlong]
        L.mcv_rig_set_input_channels.argtypes = [vp, i]
        L.mcv_orb_set_input_channels.argtypes = [vp, i]
        L.mcv_rig_sync.argtypes = [vp]
        L.mcv_rig_join.argtypes = [vp]
        L.mcv_rig_last_launches.argtypes = [vp]
        L.mcv_rig_set_chunk_frames.argtypes = [vp, i]
        L.mcv_rig_set_profiling.argtypes = [vp, i]
        L.mcv_rig_stage_ms.argtypes = [vp, vp, i, C.POINTER(i)]
        L.mcv_stereo_match.argtypes = [vp, vp, vp, vp, i, vp, vp, i, f, f, vp, vp, vp, vp]
        L.mcv_project_match.argtypes = [vp, vp, i, i, i, vp, i, vp, vp, vp, vp, vp, vp, i, f, vp, vp, C.POINTER(i)]
        L.mcv_fuse_match.argtypes = [vp, vp, i, i, i, vp, vp, i, vp, vp, vp, vp, vp, f, vp, vp, vp, vp, i, vp, vp, C.POINTER(i)]
        L.mcv_wnd_track.argtypes = [vp, vp, i, vp, i, vp, vp, i, i, i, vp, vp, vp, C.POINTER(i)]
        L.mcv_distinctive_descriptors.argtypes = [vp, vp, i, vp, vp, vp]
        L.mcv_lk_track.argtypes = [vp, vp, i, i, C.c_size_t, vp, i, vp, vp, vp]
        L.mcv_lk_track_batch.argtypes = [vp, vp, i, i, i, C.c_size_t, vp, vp, vp, vp, vp]
        L.mcv_kl_track.argtypes = [vp, vp, i, i, C.c_size_t, vp, i, vp, vp, C.POINTER(i)]
        L.mcv_voc_create.argtypes = [i, vp, vp, vp, vp, vp, i, i, i, i, C.POINTER(vp)]
        L.mcv_voc_destroy.argtypes = [vp]
        L.mcv_voc_destroy.restype = None
        L.mcv_bow_transform.argtypes = [vp, vp, i, i, vp, vp, vp, vp, vp, C.POINTER(i), vp, vp, vp, C.POINTER(i)]
        L.mcv_debug_sincosf.argtypes = [vp, i, vp, vp]
        L.mcv_debug_fast_atan2.argtypes = [vp, vp, i, vp]
        L.mcv_debug_level_keypoints.argtypes = [vp, i, i, i, vp, i, C.POINTER(i)]
        L.mcv_debug_download_blurred.argtypes = [vp, i, i, vp, sz]
        L.mcv_debug_octree_clocks.argtypes = [vp]
        L.mcv_debug_popc_peak.argtypes = [i, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        _lib = L
    return _lib


def _check(st):
    if st != 0:
        raise McvError(st, lib().mcv_last_error().decode(errors="replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class ORB:
    """MCVSLAM::ORB. `config` may be a dict with the extractor.yaml keys (ORBExtractor.cpp:13-17) or keyword args."""

    def __init__(self, nkeypoints=2000, scale_factor=1.2, nlevels=8, ini_th_fast=28, min_th_fast=15, device=0, stream=None,
                 config=None):
        if config is not None:
            nkeypoints = int(config["nkeypoints"]); scale_factor = float(config["scale_factor"]); nlevels = int(config["nlevels"])
            ini_th_fast = int(config["ORBextractor.iniThFAST"]); min_th_fast = int(config["ORBextractor.minThFAST"])
        self.nlevels = nlevels
        self._h = C.c_void_p()
        prm = OrbParams(nkeypoints, scale_factor, nlevels, ini_th_fast, min_th_fast)
        _check(lib().mcv_orb_create(C.byref(prm), device, stream, C.byref(self._h)))
        arrs = [np.empty(nlevels, np.float32) for _ in range(4)]
        q = np.empty(nlevels, np.int32)
        _check(lib().mcv_orb_get_scales(self._h, *[_p(a) for a in arrs], _p(q)))
        self.mvScaleFactor, self.mvInvScaleFactor, self.mvLevelSigma2, self.mvInvLevelSigma2 = arrs
        self.mnFeaturesPerLevel = q

    @staticmethod
    def parse_yaml(path):
        """The flat `key: value` subset of pyp::yaml that config/extractor.yaml uses."""
        cfg = {}
        for line in open(path):
            line = line.split("#", 1)[0].strip()
            if ":" in line:
                k, v = line.split(":", 1)
                cfg[k.strip()] = v.strip()
        return cfg

    @classmethod
    def from_yaml(cls, path, **kw):
        return cls(config=cls.parse_yaml(path), **kw)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().mcv_orb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def GetLevels(self):
        return self.nlevels

    def max_keypoints(self, n_seeds=0, w=0, h=0):
        """Size-independent bound, or the exact bound for a w x h image."""
        if w and h:
            return lib().mcv_orb_max_keypoints_for(self._h, w, h, n_seeds)
        return lib().mcv_orb_max_keypoints(self._h, n_seeds)

    def Extract(self, img, kps=None):
        """int Extract(const cv::Mat img, Keypoints& kps, Desps& desps). `kps` = pre-seeded keypoints (in/out in the
        reference). Returns (count, kps, desps); count is -1 for an empty image (ORBextractor.cc:834)."""
        if img is None or img.size == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        if img.ndim == 3:   # CV_8UC3 BGR: System::Track's cvtColor (src/System.cpp:60-64) runs on the device
            assert img.dtype == np.uint8 and img.shape[2] == 3
            img = np.ascontiguousarray(img)
            _check(lib().mcv_orb_set_input_channels(self._h, 3))
        else:
            assert img.dtype == np.uint8 and img.ndim == 2, "CV_8UC1 expected (ORBextractor.cc:837)"
            _check(lib().mcv_orb_set_input_channels(self._h, 1))
            if img.strides[1] != 1:
                img = np.ascontiguousarray(img)
        seeds = None if kps is None or len(kps) == 0 else np.ascontiguousarray(kps, KP_DTYPE)
        ns = 0 if seeds is None else len(seeds)
        cap = self.max_keypoints(ns, img.shape[1], img.shape[0])
        out_k = np.zeros(cap, KP_DTYPE); out_d = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        _check(lib().mcv_orb_extract(self._h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(seeds), ns, _p(out_k), _p(out_d),
                                     cap, C.byref(n)))
        return n.value, out_k[:n.value].copy(), out_d[:n.value].copy()

    def ExtractBatch(self, imgs):
        """imgs: (n, H, W) u8 host array. Returns list of (kps, desps)."""
        imgs = _u8(imgs)
        n, h, w = imgs.shape
        cap = self.max_keypoints(0, w, h)
        out_k = np.zeros((n, cap), KP_DTYPE); out_d = np.zeros((n, cap, 32), np.uint8); cnt = np.zeros(n, np.int32)
        _check(lib().mcv_orb_extract_batch(self._h, _p(imgs), n, w, h, 0, _p(out_k), _p(out_d), _p(cnt), cap, 0))
        return [(out_k[i, :cnt[i]].copy(), out_d[i, :cnt[i]].copy()) for i in range(n)]

    def mvImagePyramid(self, level, image_index=0):
        w, h = C.c_int(), C.c_int()
        _check(lib().mcv_orb_download_level(self._h, image_index, level, None, 0, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        _check(lib().mcv_orb_download_level(self._h, image_index, level, _p(out), w.value, C.byref(w), C.byref(h)))
        return out

    def debug_blurred(self, level, image_index=0):
        out = np.empty_like(self.mvImagePyramid(level, image_index))
        _check(lib().mcv_debug_download_blurred(self._h, image_index, level, _p(out), out.shape[1]))
        return out

    def debug_level_keypoints(self, level, which, image_index=0):
        cap = 1 << 17
        out = np.zeros(cap, KP_DTYPE)
        n = C.c_int(0)
        _check(lib().mcv_debug_level_keypoints(self._h, image_index, level, which, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def DistributeOctTree(self, kps, min_x, max_x, min_y, max_y, n_features):
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        cap = len(kps) + 16
        out = np.zeros(cap, KP_DTYPE)
        n = C.c_int(0)
        _check(lib().mcv_orb_distribute_octree(self._h, _p(kps), len(kps), min_x, max_x, min_y, max_y, n_features, _p(out), cap, C.byref(n)))
        return out[:n.value].copy()


class MatchRes:
    """MatchRes : std::vector<cv::DMatch> with the chainable filters (src/Matcher.cpp:23-91)."""

    def __init__(self, matches):
        self.m = np.ascontiguousarray(matches, DM_DTYPE).copy()

    def __len__(self):
        return len(self.m)

    def FilterThreshold(self, thres_hold=ORB_GOOD_THRESHOLD):
        n = C.c_int(len(self.m))
        _check(lib().mcv_filter_threshold(_p(self.m), C.byref(n), int(thres_hold)))  # int parameter truncates, as in C++
        self.m = self.m[:n.value].copy()
        return self

    def FilterOrientation(self, kps1, kps2):
        kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
        n = C.c_int(len(self.m))
        _check(lib().mcv_filter_orientation(_p(self.m), C.byref(n), _p(kps1), len(kps1), _p(kps2), len(kps2)))
        self.m = self.m[:n.value].copy()
        return self

    def FilterFMatrix(self, kps1, kps2, F12, level_sigma2):
        kps1 = np.ascontiguousarray(kps1, KP_DTYPE); kps2 = np.ascontiguousarray(kps2, KP_DTYPE)
        F = np.ascontiguousarray(F12, np.float32); ls = np.ascontiguousarray(level_sigma2, np.float32)
        n = C.c_int(len(self.m))
        _check(lib().mcv_filter_fmatrix(_p(self.m), C.byref(n), _p(kps1), len(kps1), _p(kps2), len(kps2), _p(F), _p(ls), len(ls)))
        self.m = self.m[:n.value].copy()
        return self


class MatchResKnn:
    """MatchResKnn : std::vector<std::vector<cv::DMatch>>; rows have `per` entries."""

    def __init__(self, knn):
        self.knn = np.ascontiguousarray(knn, DM_DTYPE)

    def __len__(self):
        return len(self.knn)

    def FilterRatio(self, ratio=0.6):
        nq, per = self.knn.shape
        out = np.zeros(max(nq, 1), DM_DTYPE)
        n = C.c_int(0)
        _check(lib().mcv_filter_ratio(_p(self.knn), nq, per, ratio, _p(out), C.byref(n)))
        return MatchRes(out[:n.value])


class Matcher:
    """MCVSLAM::Matcher static methods."""

    @staticmethod
    def KnnMatch(desp1, desp2, k=2):
        """KnnMatch(const cv::Mat&, const cv::Mat&, k=2) == cv::BFMatcher(NORM_HAMMING).knnMatch (src/Matcher.cpp:304-308)."""
        assert k == 2
        q = _u8(desp1); t = _u8(desp2)
        out = np.zeros((len(q), 2), DM_DTYPE)
        kk = C.c_int(0)
        _check(lib().mcv_knn2_bf(_p(q), len(q), _p(t), len(t), _p(out), C.byref(kk)))
        return MatchResKnn(out[:, :kk.value])

    KnnMatch_cv = KnnMatch

    @staticmethod
    def KnnMatchRows(desp1_rows, desp2_rows, k=2):
        """KnnMatch(const std::vector<cv::Mat>&, const std::vector<cv::Mat>&, k=2) (src/Matcher.cpp:283-302): always two
        entries per query, (0, 999) padding."""
        q = _u8(desp1_rows); t = _u8(desp2_rows)
        out = np.zeros((len(q), 2), DM_DTYPE)
        _check(lib().mcv_knn2_firstparty(_p(q), len(q), _p(t), len(t), _p(out)))
        return MatchResKnn(out)

    @staticmethod
    def KnnMatchCandidates(desp1, desp2, cand_off, cand_idx):
        q = _u8(desp1); t = _u8(desp2)
        co = np.ascontiguousarray(cand_off, np.int32); ci = np.ascontiguousarray(cand_idx, np.int32)
        out = np.zeros((len(q), 2), DM_DTYPE)
        _check(lib().mcv_knn2_candidates(_p(q), len(q), _p(t), len(t), _p(co), _p(ci), _p(out)))
        return MatchResKnn(out)

    @staticmethod
    def BFMatch(desp1, desp2):
        q = _u8(desp1); t = _u8(desp2)
        out = np.zeros(len(q), DM_DTYPE)
        _check(lib().mcv_bf_match(_p(q), len(q), _p(t), len(t), _p(out)))
        return MatchRes(out)

    @staticmethod
    def DBowMatch(desp1, bow_feat1, desp2, bow_feat2):
        """bow_feat*: dict node_id -> list of feature indices (DBoW3::FeatureVector), or the flattened triple
        (fv_nodes, fv_off, fv_idx) that Vocabulary.transform returns."""
        def flat(fv):
            if isinstance(fv, tuple):
                return np.ascontiguousarray(fv[0], np.uint32), np.ascontiguousarray(fv[1], np.int32), np.ascontiguousarray(fv[2], np.int32)
            ids = np.array(sorted(fv.keys()), np.uint32)
            off = np.zeros(len(ids) + 1, np.int32)
            idx = []
            for k, nid in enumerate(ids):
                idx.extend(fv[int(nid)]); off[k + 1] = len(idx)
            return ids, off, np.array(idx, np.int32)
        d1 = _u8(desp1); d2 = _u8(desp2)
        i1, o1, x1 = flat(bow_feat1); i2, o2, x2 = flat(bow_feat2)
        out = np.zeros((max(len(d1), 1), 2), DM_DTYPE)
        n = C.c_int(0)
        _check(lib().mcv_dbow_match(_p(d1), len(d1), _p(i1), _p(o1), _p(x1), len(i1), _p(d2), len(d2), _p(i2), _p(o2), _p(x2), len(i2),
                                    _p(out), C.byref(n)))
        return MatchResKnn(out[:n.value])


def ComputeStereoMatch(orb_left, orb_right, kps_l, desc_l, kps_r, desc_r, bf, baseline):
    """Frame::ComputeStereoMatch(LEFT, RIGHT) (src/Frame.cpp:150-328) on the pyramids held by the two extractors.
    Returns (u_right, depth_left, best_dist, best_r)."""
    kl = np.ascontiguousarray(kps_l, KP_DTYPE); kr = np.ascontiguousarray(kps_r, KP_DTYPE)
    dl = _u8(desc_l); dr = _u8(desc_r)
    n = len(kl)
    ur = np.full(n, -1, np.float32); dp = np.full(n, -1, np.float32); bd = np.full(n, -1, np.int32); br = np.full(n, -1, np.int32)
    _check(lib().mcv_stereo_match(orb_left._h, orb_right._h, _p(kl), _p(dl), n, _p(kr), _p(dr), len(kr), bf, baseline, _p(ur), _p(dp),
                                  _p(bd), _p(br)))
    return ur, dp, bd, br


def ProjectBunchMapPoints(kps, desps, w, h, scale_factors, Rcw, tcw, intrinsics, mp_xyz, mp_desc, mp_level, r_threshold=5.0):
    """Object::ProjectBunchMapPoints (src/Object.cpp:208-236) for an ordered MapPoint array. Returns (cnt, idx, dist)."""
    kps = np.ascontiguousarray(kps, KP_DTYPE); desps = _u8(desps)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    R = np.ascontiguousarray(Rcw, np.float32); t = np.ascontiguousarray(tcw, np.float32); K = np.ascontiguousarray(intrinsics, np.float32)
    xyz = np.ascontiguousarray(mp_xyz, np.float32); md = _u8(mp_desc); ml = np.ascontiguousarray(mp_level, np.int32)
    n_mp = len(ml)
    oi = np.full(n_mp, -1, np.int32); od = np.full(n_mp, -1, np.int32)
    cnt = C.c_int(0)
    _check(lib().mcv_project_match(_p(kps), _p(desps), len(kps), w, h, _p(sf), len(sf), _p(R), _p(t), _p(K), _p(xyz), _p(md), _p(ml), n_mp,
                                   r_threshold, _p(oi), _p(od), C.byref(cnt)))
    return cnt.value, oi, od


def FuseMatch(kps, desps, w, h, level_sigma2, inv_level_sigma2, Rcw, tcw, Ow, intrinsics, depth_left, bf, mp_xyz, mp_normal, mp_desc, mp_level):
    """Matching front-end of Map::Fuse (src/Map.cpp:478-527) for an ordered MapPoint array. Returns (cnt, idx, dist)."""
    kps = np.ascontiguousarray(kps, KP_DTYPE); desps = _u8(desps)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    s2, is2, R, t, O, K, dl, xyz, nrm = map(f32, (level_sigma2, inv_level_sigma2, Rcw, tcw, Ow, intrinsics, depth_left, mp_xyz, mp_normal))
    md = _u8(mp_desc); ml = np.ascontiguousarray(mp_level, np.int32)
    n_mp = len(ml)
    oi = np.full(n_mp, -1, np.int32); od = np.full(n_mp, -1, np.int32)
    cnt = C.c_int(0)
    _check(lib().mcv_fuse_match(_p(kps), _p(desps), len(kps), w, h, _p(s2), _p(is2), len(s2), _p(R), _p(t), _p(O), _p(K), _p(dl), bf, _p(xyz),
                                _p(nrm), _p(md), _p(ml), n_mp, _p(oi), _p(od), C.byref(cnt)))
    return cnt.value, oi, od


def WndTrack(kps1, desps1, q_idx, kps2, desps2, w, h):
    """Tracker::Wnd_Track (src/Tracker.cpp:341-360). Returns (cnt, idx [what the reference reports: the window's first
    candidate], best [the candidate the match belongs to], dist)."""
    kps1 = np.ascontiguousarray(kps1, KP_DTYPE); desps1 = _u8(desps1); kps2 = np.ascontiguousarray(kps2, KP_DTYPE); desps2 = _u8(desps2)
    q = np.ascontiguousarray(q_idx, np.int32)
    oi = np.full(len(q), -1, np.int32); ob = np.full(len(q), -1, np.int32); od = np.full(len(q), -1, np.int32)
    cnt = C.c_int(0)
    _check(lib().mcv_wnd_track(_p(kps1), _p(desps1), len(kps1), _p(q), len(q), _p(kps2), _p(desps2), len(kps2), w, h, _p(oi), _p(ob), _p(od),
                               C.byref(cnt)))
    return cnt.value, oi, ob, od


def ComputeDistinctiveDescriptors(desps, off):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cpp:101-150) for a batch of MapPoints: rows [off[m], off[m+1]) of
    `desps` are the descriptors point m was observed with. Returns (best_idx, best_median, desp [n_mp, 32]); best_idx = -1 and a
    zero row where a point has no observation (the reference leaves such a point untouched)."""
    desps = _u8(desps); off = np.ascontiguousarray(off, np.int32)
    n_mp = len(off) - 1
    bi = np.full(n_mp, -1, np.int32); bm = np.full(n_mp, -1, np.int32); od = np.zeros((n_mp, 32), np.uint8)
    _check(lib().mcv_distinctive_descriptors(_p(desps), _p(off), n_mp, _p(bi), _p(bm), _p(od)))
    return bi, bm, od


def LkTrack(prev, nxt, pts):
    """cv::calcOpticalFlowPyrLK as KL_Track calls it (src/Frame.cpp:52-54). Returns (next_pts [n, 2], status, err)."""
    prev = _u8(prev); nxt = _u8(nxt)
    assert prev.shape == nxt.shape and prev.ndim == 2
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2); n = len(pts)
    out = np.zeros((n, 2), np.float32); st = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
    _check(lib().mcv_lk_track(_p(prev), _p(nxt), prev.shape[1], prev.shape[0], prev.strides[0], _p(pts), n, _p(out), _p(st), _p(err)))
    return out, st, err


def LkTrackBatch(prev, nxt, pts_list):
    """mcv_lk_track_batch: prev / nxt [n_pairs, h, w] u8, pts_list = one [n_k, 2] array per pair. Returns lists (next_pts, status, err)."""
    prev = _u8(prev); nxt = _u8(nxt)
    assert prev.shape == nxt.shape and prev.ndim == 3 and len(pts_list) == prev.shape[0]
    off = np.zeros(len(pts_list) + 1, np.int32)
    off[1:] = np.cumsum([len(p) for p in pts_list])
    pts = np.ascontiguousarray(np.concatenate([np.asarray(p, np.float32).reshape(-1, 2) for p in pts_list]) if off[-1] else np.zeros((0, 2), np.float32))
    n = int(off[-1])
    out = np.zeros((n, 2), np.float32); st = np.zeros(n, np.uint8); err = np.zeros(n, np.float32)
    _check(lib().mcv_lk_track_batch(_p(prev), _p(nxt), prev.shape[0], prev.shape[2], prev.shape[1], prev.strides[1], _p(pts), _p(off), _p(out), _p(st), _p(err)))
    sl = [slice(off[k], off[k + 1]) for k in range(len(pts_list))]
    return [out[s] for s in sl], [st[s] for s in sl], [err[s] for s in sl]


def KL_Track(prev, nxt, kps):
    """KL_Track (src/Frame.cpp:34-76) without the MapPoint map. Returns (cnt, new_kps, ok)."""
    prev = _u8(prev); nxt = _u8(nxt)
    assert prev.shape == nxt.shape and prev.ndim == 2
    kps = np.ascontiguousarray(kps, KP_DTYPE); n = len(kps)
    new = np.zeros(n, KP_DTYPE); ok = np.zeros(n, np.uint8); cnt = C.c_int(0)
    _check(lib().mcv_kl_track(_p(prev), _p(nxt), prev.shape[1], prev.shape[0], prev.strides[0], _p(kps), n, _p(new), _p(ok), C.byref(cnt)))
    return cnt.value, new, ok


class Vocabulary:
    """DBoW3::Vocabulary reduced to what Object::ComputeBow needs (src/Object.cpp:238-247): transform(). `voc` = the flat arrays
    (child_off, child_ids, node_desc, word_id, weight, L, weighting, norm) described in include/mcv_b200.h."""

    def __init__(self, voc, device=0):
        co = np.ascontiguousarray(voc["child_off"], np.int32); ci = np.ascontiguousarray(voc["child_ids"], np.uint32)
        nd = _u8(voc["node_desc"]); wi = np.ascontiguousarray(voc["word_id"], np.int32); ww = np.ascontiguousarray(voc["weight"], np.float64)
        self._v = C.c_void_p()
        _check(lib().mcv_voc_create(len(wi), _p(co), _p(ci), _p(nd), _p(wi), _p(ww), int(voc["L"]), int(voc["weighting"]), int(voc["norm"]), device,
                                    C.byref(self._v)))

    def close(self):
        if getattr(self, "_v", None) and self._v.value:
            lib().mcv_voc_destroy(self._v)
            self._v = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform(self, desps, levelsup=4):
        """voc.transform(v_desps, bow_vector, bow_feature, levelsup). Returns dict(word, weight, nid, bow_ids, bow_vals, fv_nodes,
        fv_off, fv_idx) — fv_* is the FeatureVector in the flattened form Matcher.DBowMatch takes."""
        d = _u8(desps); n = len(d)
        ow = np.zeros(n, np.int32); owt = np.zeros(n, np.float64); onid = np.zeros(n, np.uint32)
        bi = np.zeros(n + 1, np.uint32); bv = np.zeros(n + 1, np.float64); fn = np.zeros(n + 1, np.uint32); fo = np.zeros(n + 2, np.int32); fi = np.zeros(n + 1, np.int32)
        nb, nf = C.c_int(0), C.c_int(0)
        _check(lib().mcv_bow_transform(self._v, _p(d), n, levelsup, _p(ow), _p(owt), _p(onid), _p(bi), _p(bv), C.byref(nb), _p(fn), _p(fo), _p(fi), C.byref(nf)))
        k, m = nb.value, nf.value
        return dict(word=ow, weight=owt, nid=onid, bow_ids=bi[:k].copy(), bow_vals=bv[:k].copy(), fv_nodes=fn[:m].copy(), fv_off=fo[:m + 1].copy(),
                    fv_idx=fi[:fo[m]].copy())


class Rig:
    """The Frame constructor's ORBE + SMatch stages for batches of (left, right, wide) triplets (src/Frame.cpp:118-137)."""

    def __init__(self, nkeypoints=2000, scale_factor=1.2, nlevels=8, ini_th_fast=28, min_th_fast=15, bf=955.40503, baseline=1.0,
                 device=0, stream=None):
        prm = RigParams(OrbParams(nkeypoints, scale_factor, nlevels, ini_th_fast, min_th_fast), bf, baseline)
        self._r = C.c_void_p()
        _check(lib().mcv_rig_create(C.byref(prm), device, stream, C.byref(self._r)))
        self.cap = lib().mcv_rig_max_keypoints(self._r)

    def close(self):
        if getattr(self, "_r", None) and self._r.value:
            lib().mcv_rig_destroy(self._r)
            self._r = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_input_channels(self, channels):
        """1 = gray (default), 3 = interleaved BGR: System::Track's cvtColor(BGR2GRAY) (src/System.cpp:60-64) runs on the device."""
        _check(lib().mcv_rig_set_input_channels(self._r, int(channels)))

    def process(self, imgs):
        """imgs: (n_frames, 3, H, W) u8 host array — or (n_frames, 3, H, W, 3) BGR after set_input_channels(3).
        Returns dict of host arrays (kps, desc, counts, u_right, depth_left)."""
        imgs = _u8(imgs)
        n, three, h, w = imgs.shape[:4]
        assert three == 3
        cap = self.cap
        out = dict(kps=np.zeros((n, 3, cap), KP_DTYPE), desc=np.zeros((n, 3, cap, 32), np.uint8), counts=np.zeros((n, 3), np.int32),
                   u_right=np.zeros((n, cap), np.float32), depth_left=np.zeros((n, cap), np.float32))
        _check(lib().mcv_rig_process(self._r, _p(imgs), n, w, h, 0, _p(out["kps"]), _p(out["desc"]), _p(out["counts"]), _p(out["u_right"]),
                                     _p(out["depth_left"]), cap, 0))
        return out

    def process_ptrs(self, imgs_ptr, n_frames, w, h, kps_ptr, desc_ptr, counts_ptr, ur_ptr, dp_ptr, on_device):
        """Raw-pointer variant (pinned host or device memory owned by the caller, e.g. torch tensors)."""
        _check(lib().mcv_rig_process(self._r, imgs_ptr, n_frames, w, h, int(on_device), kps_ptr, desc_ptr, counts_ptr, ur_ptr, dp_ptr,
                                     self.cap, int(on_device)))

    def process_async(self, imgs_ptr, n_frames, w, h, kps_ptr, desc_ptr, counts_ptr, ur_ptr, dp_ptr):
        _check(lib().mcv_rig_process_async(self._r, imgs_ptr, n_frames, w, h, kps_ptr, desc_ptr, counts_ptr, ur_ptr, dp_ptr, self.cap))

    def submit(self, imgs_ptr, n_frames, w, h, kps_ptr, desc_ptr, counts_ptr, ur_ptr, dp_ptr):
        """Enqueues one step on HOST (pinned) buffers — H2D, kernels, D2H — and returns a ticket for wait()."""
        t = C.c_longlong(0)
        _check(lib().mcv_rig_submit(self._r, imgs_ptr, n_frames, w, h, kps_ptr, desc_ptr, counts_ptr, ur_ptr, dp_ptr, self.cap, C.byref(t)))
        return t.value

    def wait(self, ticket):
        _check(lib().mcv_rig_wait(self._r, ticket))

    def join(self):
        """Orders the rig's stream after everything process_async has enqueued (no host wait)."""
        _check(lib().mcv_rig_join(self._r))

    def sync(self):
        _check(lib().mcv_rig_sync(self._r))

    def set_chunk_frames(self, n):
        _check(lib().mcv_rig_set_chunk_frames(self._r, int(n)))

    def last_launches(self):
        return lib().mcv_rig_last_launches(self._r)

    STAGES = ("pyramid", "blur", "fast_score", "nms_cells", "quadtree", "orient_desc", "stereo_match", "stereo_median")

    def set_profiling(self, on):
        _check(lib().mcv_rig_set_profiling(self._r, int(on)))

    def stage_ms(self):
        """Returns ({stage: total ms}, n_calls) since profiling was switched on."""
        ms = np.zeros(len(self.STAGES), np.float32)
        n = C.c_int(0)
        _check(lib().mcv_rig_stage_ms(self._r, _p(ms), len(ms), C.byref(n)))
        return dict(zip(self.STAGES, ms.tolist())), n.value


def debug_sincosf(a):
    a = np.ascontiguousarray(a, np.float32)
    s = np.empty_like(a); c = np.empty_like(a)
    _check(lib().mcv_debug_sincosf(_p(a), a.size, _p(s), _p(c)))
    return s, c


def debug_fast_atan2(y, x):
    y = np.ascontiguousarray(y, np.float32); x = np.ascontiguousarray(x, np.float32)
    o = np.empty_like(y)
    _check(lib().mcv_debug_fast_atan2(_p(y), _p(x), y.size, _p(o)))
    return o


def octree_clocks():
    c = np.zeros(8, np.int64)
    _check(lib().mcv_debug_octree_clocks(_p(c)))
    return c


def popc_peak(iters=4096):
    r = C.c_double(0); ms = C.c_double(0)
    _check(lib().mcv_debug_popc_peak(iters, C.byref(r), C.byref(ms)))
    return r.value, ms.value
