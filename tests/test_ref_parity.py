"""CPU: the oracle restatement (oracle/orb_oracle.cpp) against oracle/_ref — the REFERENCE'S OWN translation units
(ORBextractor.cc, ORBExtractor.cpp, Matcher.cpp, Frame.cpp, Object.cpp, MapPoint.cpp, Map.cpp, Tracker.cpp, DBoW3) compiled
unmodified by oracle/build_ref.py against oracle/ref_shim. This is the pin SURVEY.md §8(c) asks for: every first-party block
of the path (cell loop, quadtree, IC_Angle, rBRIEF, operator() assembly, the Matcher statics, the four filters,
ComputeStereoMatch, the 30x30 grid + ProjectBunchMapPoints, Fuse, Wnd_Track, ComputeBow, ComputeDistinctiveDescriptors,
KL_Track) is executed from the reference's code and compared bit for bit. The OpenCV primitives underneath _ref are the
cv2-pinned models (tests/test_oracle_golden.py::test_primitives_vs_cv2 keeps them equal to the real cv2 4.13)."""
import os

import numpy as np
import pytest

from mcvslam_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    from oracle import ref as R
    if not R.available():
        pytest.skip("oracle/_ref/libmcv_ref.so absent and /root/reference not present to build it")
    R.lib()
    return R


def _same_kd(a, b):
    (na, ka, da), (nb, kb, db) = a, b
    assert na == nb
    assert ka.tobytes() == kb.tobytes()
    assert da.tobytes() == db.tobytes()


def edge_images():
    rng = np.random.default_rng(77)
    flat = np.full((120, 160), 90, np.uint8)
    noise = rng.integers(0, 256, (97, 131), dtype=np.uint8)
    salt = np.full((200, 240), 20, np.uint8); salt[rng.integers(0, 200, 300), rng.integers(0, 240, 300)] = 250
    yy, xx = np.mgrid[0:150, 0:210]
    checker = (((yy // 9) + (xx // 9)) % 2 * 200 + 20).astype(np.uint8)
    return dict(flat=flat, noise=noise, salt=salt, checker=checker)


# ---- A0-A9 --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [(2000, 1.2, 8, 28, 15), (2000, 1.2, 1, 28, 15), (300, 1.2, 4, 28, 15), (5000, 1.2, 8, 28, 15), (500, 1.5, 3, 40, 7)])
def test_init_parameters(oracle, ref, cfg):
    a, b = oracle.Orb(*cfg), ref.Orb(*cfg)
    for f in ("scale", "inv_scale", "sigma2", "inv_sigma2", "quota", "umax"):
        assert getattr(a, f).tobytes() == getattr(b, f).tobytes(), f


def test_extract_equals_reference(oracle, ref, golden):
    O, R = oracle.Orb(2000, 1.2, 8, 28, 15), ref.Orb(2000, 1.2, 8, 28, 15)
    for i, s in enumerate((1000, 1001, 7, 58)):
        img = synth.scene(s)
        ro = O.extract(img); rr = R.extract(img)
        _same_kd(ro, rr)
        for l in range(8):
            assert np.array_equal(O.level(l), R.level(l))
        if i < 2:                                    # ... and both equal the cv2-generated fixture
            assert rr[1].tobytes() == golden[f"g1_kps{i}"].tobytes() and rr[2].tobytes() == golden[f"g1_desc{i}"].tobytes()
    img = synth.scene(42, 512, 512)                  # shipped config: nlevels 1
    _same_kd(oracle.Orb(2000, 1.2, 1, 28, 15).extract(img), ref.Orb(2000, 1.2, 1, 28, 15).extract(img))
    _same_kd(oracle.Orb(300, 1.2, 4, 28, 15).extract(golden["g3_img"]), ref.Orb(300, 1.2, 4, 28, 15).extract(golden["g3_img"]))
    img = synth.scene(5, 1280, 720)                  # configs[3] geometry
    _same_kd(oracle.Orb(5000, 1.2, 8, 28, 15).extract(img), ref.Orb(5000, 1.2, 8, 28, 15).extract(img))


def _ref_defined(oracle, img, cfg):
    """The reference's DistributeOctTree calls q_nodes.top() on an EMPTY heap when a level has no FAST candidate at all
    (ORBextractor.cc:557-559) and divides by nCols / nRows == 0 on levels narrower than 35 px of usable area (:599-602): undefined
    behaviour (a crash here). The oracle and the engine define those levels as 'no keypoints'; _ref is only run where it is defined."""
    E = oracle.Orb(*cfg, debug=True)
    r = E.extract(img)
    if r[0] < 0:
        return False, r
    return all(len(E.debug_kps(0, l)) > 0 for l in range(cfg[2])), r


def test_extract_edge_images_and_triplets(oracle, ref):
    n_cmp = 0
    for name, img in edge_images().items():
        for cfg in ((500, 1.2, 4, 28, 15), (2000, 1.2, 3, 20, 7), (100, 1.3, 2, 28, 15)):
            ok, ro = _ref_defined(oracle, img, cfg)
            if ok:
                _same_kd(ro, ref.Orb(*cfg).extract(img)); n_cmp += 1
    assert n_cmp >= 5
    O, R = oracle.Orb(), ref.Orb()
    for s in (21, 22):
        for img in synth.triplet(s):
            _same_kd(O.extract(img), R.extract(img))
    for (w, h) in ((333, 217), (211, 187), (1000, 200)):          # odd sizes; aspect > 4.5 gives 5 quadtree roots
        img = synth.scene(9, w, h)
        ok, ro = _ref_defined(oracle, img, (800, 1.2, 5, 28, 15))
        assert ok
        _same_kd(ro, ref.Orb(800, 1.2, 5, 28, 15).extract(img))


def test_extract_empty_and_seeds(oracle, ref):
    R = ref.Orb()
    assert R.extract(np.zeros((0, 0), np.uint8))[0] == -1                       # ORBextractor.cc:834
    img = synth.scene(31)
    O = oracle.Orb()
    n, k, d = O.extract(img)
    seeds = k[[3, 50, 700, 1500]].copy()
    seeds["octave"] = [0, 0, 2, 7]
    seeds["x"] += np.float32(0.37); seeds["y"] -= np.float32(0.21)
    for i, s in enumerate(seeds):                                                 # keep them inside their level's usable area
        sc = O.scale[s["octave"]]
        seeds["x"][i] = np.float32(min(max(s["x"] / sc, 30), 100)); seeds["y"][i] = np.float32(min(max(s["y"] / sc, 30), 90))
    _same_kd(O.extract(img, seeds), R.extract(img, seeds))


def test_distribute_octree_equals_reference(oracle, ref):
    rng = np.random.default_rng(4)
    for case in range(40):
        w = int(rng.integers(60, 900)); h = int(rng.integers(40, 500))
        if h > w:
            w, h = h, w                                    # nIni = round(w/h) must be >= 1 (the reference divides by it)
        n = int(rng.integers(1, 3000)); N = int(rng.integers(1, 1200))
        if case % 4 == 0:                                  # clustered: deep trees, many equal-size ties
            cx = rng.integers(0, w, 6); cy = rng.integers(0, h, 6); c = rng.integers(0, 6, n)
            x = np.clip(cx[c] + rng.integers(-8, 9, n), 0, w - 1); y = np.clip(cy[c] + rng.integers(-8, 9, n), 0, h - 1)
        else:
            x = rng.integers(0, w, n); y = rng.integers(0, h, n)
        # distinct pixels only, in row-major candidate order, as FAST + NMS delivers them: two candidates on one pixel can never be
        # separated by DivideNode and the reference's loop (ORBextractor.cc:557-565) would not terminate
        code = np.unique(y * w + x)
        if case % 5 != 0:
            code = rng.permutation(code)
        n = len(code)
        k = np.zeros(n, oracle.KP_DTYPE)
        k["x"] = code % w; k["y"] = code // w
        k["response"] = rng.integers(15, 40 if case % 3 == 0 else 255, n)
        k["size"] = 7; k["angle"] = -1; k["class_id"] = -1
        a = oracle.distribute_octree(k, 0, w, 0, h, N); b = ref.distribute_octree(k, 0, w, 0, h, N)
        assert a.tobytes() == b.tobytes(), case


# ---- A10-A12 ------------------------------------------------------------------------------------------------------------
def test_matcher_family_equals_reference(oracle, ref, golden):
    pop = np.array([bin(i).count("1") for i in range(256)])
    a = synth.descriptors(64, 1); b = synth.descriptors(64, 2)
    for i in range(64):
        assert ref.hamming(a[i], b[i]) == int(pop[a[i] ^ b[i]].sum())
    cases = [(synth.descriptors(300, 1), synth.descriptors(400, 2)), (synth.descriptors(257, 3, True), synth.descriptors(513, 4, True)),
             (golden["g4_q"], golden["g4_t"]), (synth.descriptors(5, 5), synth.descriptors(1, 6)), (synth.descriptors(1, 7), synth.descriptors(2001, 8))]
    for q, t in cases:
        assert oracle.knn2_firstparty(q, t).tobytes() == ref.knn2_firstparty(q, t).tobytes()
        ro, ko = oracle.knn2_bf(q, t)
        for cv_variant in (False, True):
            rr, kr = ref.knn2_bf(q, t, cv_variant)
            assert ko == kr
            assert ro[:, :ko].tobytes() == rr[:, :kr].tobytes()
        m = ref.bf_match(q, t)
        assert np.array_equal(m["trainIdx"], ro[:, 0]["trainIdx"]) and np.array_equal(m["distance"], ro[:, 0]["distance"])
    q = synth.descriptors(0, 1)
    assert oracle.knn2_firstparty(synth.descriptors(3, 1), q).tobytes() == ref.knn2_firstparty(synth.descriptors(3, 1), q).tobytes()
    rng = np.random.default_rng(3)
    q = synth.descriptors(500, 11, True); t = synth.descriptors(700, 12, True)
    lens = rng.integers(0, 9, 500); lens[:20] = 0; lens[20:40] = 1
    off = np.zeros(501, np.int32); off[1:] = np.cumsum(lens)
    ci = rng.integers(0, 700, off[-1]).astype(np.int32)
    assert oracle.knn2_candidates(q, t, off, ci).tobytes() == ref.knn2_candidates(q, t, off, ci).tobytes()


def _matches(oracle, n1=1500, n2=1600, n=900, seed=0):
    rng = np.random.default_rng(seed)
    k1 = np.zeros(n1, oracle.KP_DTYPE); k2 = np.zeros(n2, oracle.KP_DTYPE)
    for k in (k1, k2):
        k["x"] = rng.uniform(0, 640, len(k)).astype(np.float32); k["y"] = rng.uniform(0, 480, len(k)).astype(np.float32)
        k["angle"] = rng.uniform(0, 360, len(k)).astype(np.float32); k["octave"] = rng.integers(0, 8, len(k))
    m = np.zeros(n, oracle.DM_DTYPE)
    m["queryIdx"] = rng.integers(0, n1, n); m["trainIdx"] = rng.integers(0, n2, n); m["distance"] = rng.integers(0, 100, n)
    return k1, k2, m


def test_filters_equal_reference(oracle, ref):
    rng = np.random.default_rng(9)
    knn = np.zeros((400, 2), oracle.DM_DTYPE)
    knn["distance"] = rng.integers(0, 120, (400, 2)); knn["distance"][:30] = 0; knn["distance"][30:60, 1] = 999
    knn["queryIdx"] = np.arange(400)[:, None]; knn["trainIdx"] = rng.integers(0, 999, (400, 2))
    for r in (0.6, 0.7, 0.75):
        assert oracle.filter_ratio(knn, r).tobytes() == ref.filter_ratio(knn, r).tobytes()
    assert oracle.filter_ratio(knn[:, :1], 0.6).tobytes() == ref.filter_ratio(knn[:, :1], 0.6).tobytes()
    for seed in range(4):
        k1, k2, m = _matches(oracle, seed=seed)
        for th in (46, 34, 0, 200):
            assert oracle.filter_threshold(m, th).tobytes() == ref.filter_threshold(m, th).tobytes()
        k2b = k2.copy()
        if seed == 1:                                   # concentrated rotations: bins with equal sizes exercise std::sort's order
            k2b["angle"] = (k1["angle"][m["queryIdx"] % len(k1)][:len(k2)] if False else k2["angle"])
            k2b["angle"] = np.float32(10.0); k1 = k1.copy(); k1["angle"] = (rng.integers(0, 6, len(k1)) * 9 + 10).astype(np.float32)
        assert oracle.filter_orientation(m, k1, k2b).tobytes() == ref.filter_orientation(m, k1, k2b).tobytes()
    assert len(ref.filter_orientation(m[:0], k1, k2)) == 0 and len(oracle.filter_orientation(m[:0], k1, k2)) == 0


def test_filter_fmatrix_equals_reference(oracle, ref):
    """MatchRes::FilterFMatrix / CheckDistEpipolarLine (src/Matcher.cpp:76-91,310-325): random F, den == 0, swap-remove order."""
    s2 = oracle.Orb().sigma2
    rng = np.random.default_rng(12)
    for seed in range(6):
        k1, k2, m = _matches(oracle, seed=seed + 10)
        if seed < 3:
            # a real epipolar geometry (pure sideways translation: F = [t]x) so that a good share of the matches survives
            F = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32) * np.float32(rng.uniform(0.5, 2))
            k2 = k2.copy(); sel = rng.random(len(m)) < 0.5
            k2["y"][m["trainIdx"][sel]] = k1["y"][m["queryIdx"][sel]] + rng.normal(0, 1.5, int(sel.sum())).astype(np.float32)
        elif seed == 3:
            F = np.zeros((3, 3), np.float32)            # a = b = 0: den == 0 rejects everything
        else:
            F = rng.normal(0, 1, (3, 3)).astype(np.float32)
        a = oracle.filter_fmatrix(m, k1, k2, F, s2); b = ref.filter_fmatrix(m, k1, k2, F, s2)
        assert a.tobytes() == b.tobytes(), seed
        if seed < 3:
            assert 0 < len(a) < len(m)
        if seed == 3:
            assert len(a) == 0


def oracle_dbow_match(oracle, d1, fv1, d2, fv2):
    """src/Matcher.cpp:146-193 on the oracle's candidate 2-NN: common nodes ascending, queries in the node's list order, queries
    without a second neighbour dropped; imgIdx = -1 (three-argument DMatch constructor)."""
    exp = []
    for nid in sorted(set(fv1) & set(fv2)):
        c = np.array(fv2[nid], np.int32)
        for f1 in fv1[nid]:
            r = oracle.knn2_candidates(d1[f1:f1 + 1], d2, np.array([0, len(c)], np.int32), c)[0]
            if r[1]["distance"] != 999:
                exp.append([(f1, c[r[0]["trainIdx"]], -1, r[0]["distance"]), (f1, c[r[1]["trainIdx"]], -1, r[1]["distance"])])
    return np.array(exp, oracle.DM_DTYPE).reshape(-1, 2)


def test_dbow_match_equals_reference(oracle, ref):
    rng = np.random.default_rng(8)
    d1 = synth.descriptors(600, 1, True); d2 = synth.descriptors(700, 2, True)
    nodes = np.arange(3, 60, 2)
    fv1 = {int(nd): sorted(rng.choice(600, rng.integers(1, 30), replace=False).tolist()) for nd in nodes if rng.random() < 0.8}
    fv2 = {int(nd): sorted(rng.choice(700, rng.integers(1, 30), replace=False).tolist()) for nd in nodes if rng.random() < 0.8}
    fv2[1000] = [5]; fv1[2] = [7, 9]
    r = ref.dbow_match(d1, fv1, d2, fv2)
    o = oracle_dbow_match(oracle, d1, fv1, d2, fv2)
    assert len(o) > 100 and r.tobytes() == o.tobytes()


# ---- A13 ----------------------------------------------------------------------------------------------------------------
def test_stereo_equals_reference(oracle, ref):
    rig = ref.Rig()
    for seed in (21, 23, 40):
        trip = synth.triplet(seed)
        out = rig.frame(trip)                                               # the real Frame constructor: ThreadPool(3) + SMatch
        L, R = oracle.Orb(), oracle.Orb()
        nl, kl, dl = L.extract(trip[0]); nr, kr, dr = R.extract(trip[1])
        assert out["counts"][0] == nl and out["counts"][1] == nr
        assert out["kps"][0, :nl].tobytes() == kl.tobytes() and out["desc"][1, :nr].tobytes() == dr.tobytes()
        n, ur, dp, bd, br = oracle.stereo_match(L, R, kl, dl, kr, dr, 480, 955.40503, 1.0)
        assert n > 200
        assert out["u_right"][:nl].tobytes() == ur.tobytes()
        assert out["depth_left"][:nl].tobytes() == dp.tobytes()
    # other bf / baseline (maxD clamps), and caller keypoints against the extractors' pyramids
    rig2 = ref.Rig(bf=40.0, baseline=0.5)
    trip = synth.triplet(25)
    out = rig2.frame(trip)
    L, R = oracle.Orb(), oracle.Orb()
    nl, kl, dl = L.extract(trip[0]); nr, kr, dr = R.extract(trip[1])
    n, ur, dp, bd, br = oracle.stereo_match(L, R, kl, dl, kr, dr, 480, 40.0, 0.5)
    assert out["u_right"][:nl].tobytes() == ur.tobytes() and out["depth_left"][:nl].tobytes() == dp.tobytes()


# ---- A14 + f1 -----------------------------------------------------------------------------------------------------------
def _mappoints(k, d, n_mp, rng, fx, fy, cx, cy, tt):
    n = len(k)
    src = rng.integers(0, n, n_mp)
    z = rng.uniform(2, 50, n_mp).astype(np.float32)
    u = k["x"][src] + rng.normal(0, 2.0, n_mp).astype(np.float32); v = k["y"][src] + rng.normal(0, 2.0, n_mp).astype(np.float32)
    pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1).astype(np.float32)
    pw = (pc - tt).astype(np.float32)
    md = d[src].copy()
    flips = rng.integers(0, 40, n_mp)
    for i in range(n_mp):
        b = rng.choice(256, flips[i], replace=False)
        np.bitwise_xor.at(md[i], b // 8, (1 << (b % 8)).astype(np.uint8))
    pw[::17, 2] = -pw[::17, 2]                                             # behind the camera
    return pw, md, k["octave"][src].astype(np.int32)


def test_grid_and_projection_equal_reference(oracle, ref):
    O = oracle.Orb(); R = ref.Orb()
    img = synth.scene(55)
    n, k, d = O.extract(img)
    fx = fy = np.float32(955.40503 * 640 / 512); cx, cy = np.float32(320), np.float32(240)
    Rcw = np.eye(3, dtype=np.float32); tt = np.array([0.02, -0.01, 0.03], np.float32)
    obj = ref.Obj(R, k, d, 640, 480, [fx, fy, cx, cy], Rcw, tt)
    rng = np.random.default_rng(6)
    pw, md, lvl = _mappoints(k, d, 3000, rng, fx, fy, cx, cy, tt)
    for r_th in (5.0, 7.0, 10.0):
        cr, ir = obj.project_match(pw, md, lvl, r_th)
        co, io, do = oracle.project_match(k, d, 640, 480, O.scale, Rcw, tt, [fx, fy, cx, cy], pw, md, lvl, r_th)
        assert cr == co and cr > 300
        assert np.array_equal(ir, io)


def test_fuse_and_wnd_track_equal_reference(oracle, ref):
    O = oracle.Orb(); R = ref.Orb()
    img = synth.scene(58)
    n, k, d = O.extract(img)
    fx = fy = np.float32(900.0); cx, cy = np.float32(320), np.float32(240)
    Rcw = np.eye(3, dtype=np.float32); tt = np.array([0.05, 0.02, -0.04], np.float32)
    Ow = (-Rcw.T @ tt).astype(np.float32)
    rng = np.random.default_rng(16)
    pw, md, lvl = _mappoints(k, d, 1500, rng, fx, fy, cx, cy, tt)
    view = pw - Ow
    nrm = (view / np.linalg.norm(view, axis=1, keepdims=True)).astype(np.float32)
    nrm[::13] *= -1                                                          # back-facing
    dl = np.full(n, -1, np.float32); has = rng.random(n) < 0.5
    dl[has] = rng.uniform(2, 50, int(has.sum())).astype(np.float32)
    obj = ref.Obj(R, k, d, 640, 480, [fx, fy, cx, cy], Rcw, tt)
    cr, ir = obj.fuse_match(dl, 900.0, pw, nrm, md, lvl)
    co, io, do = oracle.fuse_match(k, d, 640, 480, O.sigma2, O.inv_sigma2, Rcw, tt, Ow, [fx, fy, cx, cy], dl, 900.0, pw, nrm, md, lvl)
    assert cr == co and cr > 100
    assert np.array_equal(ir, io)
    b = np.roll(img, (3, -5), (0, 1))
    n2, k2, d2 = O.extract(b)
    qi = np.sort(rng.choice(n, 800, replace=False)).astype(np.int32)
    o1 = ref.Obj(R, k, d, 640, 480); o2 = ref.Obj(R, k2, d2, 640, 480)
    cr, ir = ref.wnd_track(o1, o2, qi)
    co, io, ob, od = oracle.wnd_track(k, d, qi, k2, d2, 640, 480)
    assert cr == co and cr > 100
    assert np.array_equal(ir, io)


# ---- f2, f4 -------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("weighting,norm", [(0, 1), (1, 2), (3, 1), (2, 0)])
def test_compute_bow_equals_reference(oracle, ref, weighting, norm, tmp_path):
    voc = synth.random_vocabulary(5 + weighting, K=7, L=4, weighting=weighting, norm=norm)
    path = ref.write_dbow3_binary(voc, str(tmp_path / "voc.dbow3"))
    assert ref.voc_load(path) > 0                                           # DBoW3's own binary loader
    d = synth.descriptors(900, 3)
    d[:200] = voc["node_desc"][np.random.default_rng(1).integers(1, len(voc["node_desc"]), 200)]
    k = np.zeros(len(d), oracle.KP_DTYPE)
    obj = ref.Obj(ref.Orb(), k, d, 640, 480)
    r = obj.compute_bow()                                                    # Object::ComputeBow: voc.transform(desps, bow, feat, 4)
    o = oracle.bow_transform(d, voc, 4)
    assert np.array_equal(r["bow_ids"], o["bow_ids"])
    assert r["bow_vals"].tobytes() == o["bow_vals"].tobytes()
    assert np.array_equal(r["fv_nodes"], o["fv_nodes"]) and np.array_equal(r["fv_off"], o["fv_off"]) and np.array_equal(r["fv_idx"], o["fv_idx"])


@pytest.mark.skipif(not os.path.exists("/root/reference/Vocabulary/orbvoc.dbow3"), reason="the shipped vocabulary lives in /root/reference")
def test_compute_bow_on_the_shipped_vocabulary(oracle, ref, tmp_path):
    """The WHOLE Vocabulary/orbvoc.dbow3 (971 814 words, quicklz-compressed) through DBoW3's own loader; the oracle gets the same
    tree as flat arrays (Vocabulary::save uncompressed -> ref.read_dbow3_binary); 2000 real ORB descriptors + random ones."""
    try:
        assert ref.voc_load("/root/reference/Vocabulary/orbvoc.dbow3") == 971814
        assert ref.voc_save_uncompressed(str(tmp_path / "full.bin")) == 0
        voc = ref.read_dbow3_binary(str(tmp_path / "full.bin"))
        assert voc["K"] == 10 and voc["L"] == 6 and voc["weighting"] == 0 and voc["norm"] == 1 and len(voc["word_id"]) == 1082073
        n, k, d = oracle.Orb(2000, 1.2, 8, 28, 15).extract(synth.scene(31))
        d = np.concatenate([d, synth.descriptors(300, 4)])
        r = ref.Obj(ref.Orb(), np.zeros(len(d), oracle.KP_DTYPE), d, 640, 480).compute_bow()
        o = oracle.bow_transform(d, voc, 4)
        assert len(r["bow_ids"]) > 1500
        assert np.array_equal(r["bow_ids"], o["bow_ids"]) and r["bow_vals"].tobytes() == o["bow_vals"].tobytes()
        assert np.array_equal(r["fv_nodes"], o["fv_nodes"]) and np.array_equal(r["fv_off"], o["fv_off"]) and np.array_equal(r["fv_idx"], o["fv_idx"])
    finally:
        small = synth.random_vocabulary(5, K=7, L=4)      # do not leave 1 M nodes in the reference's process-wide Object::voc
        ref.voc_load(ref.write_dbow3_binary(small, str(tmp_path / "small.dbow3")))


def test_distinctive_equals_reference(oracle, ref):
    from test_oracle_golden import _ragged_observations
    sizes = [0, 1, 2, 3, 4, 5, 8, 33, 64, 100, 0, 7]
    desc, off = _ragged_observations(11, sizes)
    # the reference walks an unordered_map<KeyFrame, ...>: present the observations to the oracle in that same order
    L = ref.lib()
    perm_desc = desc.copy(); orders = []
    for m, n in enumerate(sizes):
        order = np.zeros(max(n, 1), np.int32)
        L.ref_distinctive_order(n, order.ctypes.data)
        orders.append(order[:n])
        perm_desc[off[m]:off[m + 1]] = desc[off[m]:off[m + 1]][order[:n]]
    br, od = ref.distinctive(ref.Orb(), desc, off)
    bi, bm = oracle.distinctive(perm_desc, off)
    for m, n in enumerate(sizes):
        if n == 0:
            assert br[m] == -1 and bi[m] == -1
            continue
        assert od[m].tobytes() == perm_desc[off[m] + bi[m]].tobytes(), m


def test_kl_track_equals_reference(oracle, ref):
    a = oracle.gauss7(oracle.gauss7(synth.scene(58))); b = synth.shifted(a, 2.4, 1.1, 9, noise=0)   # err < 1: mean residual below one grey level
    O = oracle.Orb(); n, k, d = O.extract(a)
    k = k[::3].copy()
    cnt_r, new_r, src_r = ref.kl_track(ref.Orb(), a, b, k)
    cnt_o, new_o, ok, nxt, st, err = oracle.kl_track(a, b, k)
    assert cnt_r == cnt_o and cnt_r > 300
    # the reference appends in GetMapPointsVector() (unordered) order: compare as sets keyed by the source keypoint
    got = {int(s): new_r[j].tobytes() for j, s in enumerate(src_r)}
    want = {int(i): new_o[i].tobytes() for i in np.nonzero(ok)[0]}
    assert got == want
