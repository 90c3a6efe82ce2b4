"""GPU parity against the REFERENCE ITSELF: the CUDA path through the C ABI vs oracle/_ref/libmcv_ref.so — the reference's own
ORBextractor.cc / ORBExtractor.cpp / Matcher.cpp / Frame.cpp / Object.cpp / Map.cpp / Tracker.cpp / MapPoint.cpp / DBoW3 sources
compiled unmodified (oracle/build_ref.py; the prebuilt .so travels to the GPU box, /root/reference is not read at run time).
Bit-exact bar for every field (angles and responses included: 0 ulp)."""
import numpy as np
import pytest

from mcvslam_b200 import synth
from test_ref_parity import _mappoints, oracle_dbow_match, ref  # noqa: F401  (fixture)
from test_oracle_golden import _ragged_observations

pytestmark = pytest.mark.gpu


def _same_kd(n, k, d, rr, tag=""):
    nr, kr, dr = rr
    assert n == nr, (tag, n, nr)
    for f in kr.dtype.names:
        a, b = (k[f].view(np.uint32), kr[f].view(np.uint32)) if k[f].dtype.kind == "f" else (k[f], kr[f])
        bad = np.nonzero(a != b)[0]
        assert len(bad) == 0, f"{tag}: field {f} differs at {bad[:8]} ({len(bad)} of {n})"
    assert d.tobytes() == dr.tobytes(), tag


@pytest.mark.parametrize("cfg,shape,seeds", [((2000, 1.2, 8, 28, 15), (480, 640), (1000, 1001, 7)), ((5000, 1.2, 8, 28, 15), (720, 1280), (5,)),
                                             ((2000, 1.2, 1, 28, 15), (512, 512), (42,)), ((800, 1.2, 5, 28, 15), (217, 333), (9,)),
                                             ((800, 1.2, 5, 28, 15), (200, 1000), (9,)), ((500, 1.5, 3, 40, 7), (480, 640), (3,))])
def test_extract_vs_reference(api, ref, cfg, shape, seeds):
    E = api.ORB(*cfg); R = ref.Orb(*cfg)
    for s in seeds:
        img = synth.scene(s, shape[1], shape[0])
        n, k, d = E.Extract(img)
        _same_kd(n, k, d, R.extract(img), f"{cfg} {shape} seed {s}")
        for l in range(cfg[2]):
            assert np.array_equal(E.mvImagePyramid(l), R.level(l)), l
    assert (E.mvScaleFactor == R.scale).all() and (E.mvLevelSigma2 == R.sigma2).all() and (E.mnFeaturesPerLevel == R.quota).all()


def test_rig_frames_vs_reference_frame_ctor(api, ref):
    """mcv_rig_process == the reference's Frame constructor (ThreadPool(3) extraction + ComputeStereoMatch), frame by frame."""
    frames = np.stack([synth.triplet(s) for s in (21, 23, 40, 41, 42)])
    rig = api.Rig()
    out = rig.process(frames)
    R = ref.Rig()
    n_stereo = 0
    for f in range(len(frames)):
        r = R.frame(frames[f])
        assert np.array_equal(out["counts"][f], r["counts"])
        for c in range(3):
            n = r["counts"][c]
            assert out["kps"][f, c, :n].tobytes() == r["kps"][c, :n].tobytes(), (f, c)
            assert out["desc"][f, c, :n].tobytes() == r["desc"][c, :n].tobytes(), (f, c)
        nl = r["counts"][0]
        assert out["u_right"][f, :nl].tobytes() == r["u_right"][:nl].tobytes(), f
        assert out["depth_left"][f, :nl].tobytes() == r["depth_left"][:nl].tobytes(), f
        n_stereo += int((r["u_right"][:nl] >= 0).sum())
    assert n_stereo > 500


def test_matcher_vs_reference(api, ref, golden):
    cases = [(synth.descriptors(2000, 1), synth.descriptors(2003, 2)), (synth.descriptors(257, 3, True), synth.descriptors(4100, 4, True)),
             (golden["g4_q"], golden["g4_t"]), (synth.descriptors(5, 5), synth.descriptors(1, 6)), (synth.descriptors(1, 7), synth.descriptors(2001, 8))]
    for q, t in cases:
        rr, kr = ref.knn2_bf(q, t)
        g = api.Matcher.KnnMatch(q, t).knn
        assert g.shape[1] == kr and g.tobytes() == rr[:, :kr].tobytes()
        assert api.Matcher.KnnMatchRows(q, t).knn.tobytes() == ref.knn2_firstparty(q, t).tobytes()
        assert api.Matcher.BFMatch(q, t).m.tobytes() == ref.bf_match(q, t).tobytes()
    rng = np.random.default_rng(3)
    q = synth.descriptors(500, 11, True); t = synth.descriptors(700, 12, True)
    lens = rng.integers(0, 9, 500); lens[:20] = 0; lens[20:40] = 1
    off = np.zeros(501, np.int32); off[1:] = np.cumsum(lens)
    ci = rng.integers(0, 700, off[-1]).astype(np.int32)
    assert api.Matcher.KnnMatchCandidates(q, t, off, ci).knn.tobytes() == ref.knn2_candidates(q, t, off, ci).tobytes()
    fv1 = {int(nd): sorted(rng.choice(500, rng.integers(1, 30), replace=False).tolist()) for nd in range(3, 60, 2) if rng.random() < 0.8}
    fv2 = {int(nd): sorted(rng.choice(700, rng.integers(1, 30), replace=False).tolist()) for nd in range(3, 60, 2) if rng.random() < 0.8}
    assert api.Matcher.DBowMatch(q, fv1, t, fv2).knn.tobytes() == ref.dbow_match(q, fv1, t, fv2).tobytes()


def test_projection_fuse_wnd_vs_reference(api, ref):
    E = api.ORB(); R = ref.Orb()
    img = synth.scene(55)
    n, k, d = E.Extract(img)
    fx = fy = np.float32(955.40503 * 640 / 512); cx, cy = np.float32(320), np.float32(240)
    Rcw = np.eye(3, dtype=np.float32); tt = np.array([0.02, -0.01, 0.03], np.float32)
    K = [fx, fy, cx, cy]
    obj = ref.Obj(R, k, d, 640, 480, K, Rcw, tt)
    rng = np.random.default_rng(6)
    pw, md, lvl = _mappoints(k, d, 10000, rng, fx, fy, cx, cy, tt)                   # configs[2]: 10 k MapPoints
    for r_th in (5.0, 7.0, 10.0):
        cr, ir = obj.project_match(pw, md, lvl, r_th)
        cg, ig, dg = api.ProjectBunchMapPoints(k, d, 640, 480, E.mvScaleFactor, Rcw, tt, K, pw, md, lvl, r_th)
        assert cg == cr and cr > 1000 and np.array_equal(ig, ir)
    Ow = (-Rcw.T @ tt).astype(np.float32)
    view = pw[:3000] - Ow
    nrm = (view / np.linalg.norm(view, axis=1, keepdims=True)).astype(np.float32); nrm[::13] *= -1
    dl = np.full(n, -1, np.float32); has = rng.random(n) < 0.5
    dl[has] = rng.uniform(2, 50, int(has.sum())).astype(np.float32)
    cr, ir = obj.fuse_match(dl, 900.0, pw[:3000], nrm, md[:3000], lvl[:3000])
    cg, ig, dg = api.FuseMatch(k, d, 640, 480, E.mvLevelSigma2, E.mvInvLevelSigma2, Rcw, tt, Ow, K, dl, 900.0, pw[:3000], nrm, md[:3000], lvl[:3000])
    assert cg == cr and cr > 200 and np.array_equal(ig, ir)
    b = np.roll(img, (3, -5), (0, 1))
    n2, k2, d2 = E.Extract(b)
    qi = np.sort(rng.choice(n, 1200, replace=False)).astype(np.int32)
    cr, ir = ref.wnd_track(ref.Obj(R, k, d, 640, 480), ref.Obj(R, k2, d2, 640, 480), qi)
    cg, ig, bg, dg = api.WndTrack(k, d, qi, k2, d2, 640, 480)
    assert cg == cr and cr > 200 and np.array_equal(ig, ir)


def test_bow_distinctive_kl_vs_reference(api, ref, oracle, tmp_path):
    voc = synth.random_vocabulary(5, K=7, L=4, weighting=0, norm=1)
    assert ref.voc_load(ref.write_dbow3_binary(voc, str(tmp_path / "voc.dbow3"))) > 0
    d = synth.descriptors(900, 3)
    d[:200] = voc["node_desc"][np.random.default_rng(1).integers(1, len(voc["node_desc"]), 200)]
    r = ref.Obj(ref.Orb(), np.zeros(len(d), api.KP_DTYPE), d, 640, 480).compute_bow()
    g = api.Vocabulary(voc).transform(d, 4)
    assert np.array_equal(g["bow_ids"], r["bow_ids"]) and g["bow_vals"].tobytes() == r["bow_vals"].tobytes()
    assert np.array_equal(g["fv_nodes"], r["fv_nodes"]) and np.array_equal(g["fv_off"], r["fv_off"]) and np.array_equal(g["fv_idx"], r["fv_idx"])
    sizes = [0, 1, 2, 3, 4, 5, 8, 33, 64, 100, 0, 7]
    desc, off = _ragged_observations(11, sizes)
    L = ref.lib()
    perm = desc.copy()                                      # the reference walks an unordered_map: present its order to the engine
    for m, nobs in enumerate(sizes):
        order = np.zeros(max(nobs, 1), np.int32)
        L.ref_distinctive_order(nobs, order.ctypes.data)
        perm[off[m]:off[m + 1]] = desc[off[m]:off[m + 1]][order[:nobs]]
    br, od = ref.distinctive(ref.Orb(), desc, off)
    bi, bm, gd = api.ComputeDistinctiveDescriptors(perm, off)
    for m, nobs in enumerate(sizes):
        if nobs:
            assert gd[m].tobytes() == od[m].tobytes(), m
    a = oracle.gauss7(oracle.gauss7(synth.scene(58))); b = synth.shifted(a, 2.4, 1.1, 9, noise=0)
    n, k, dd = api.ORB().Extract(a)
    sel = k[::3].copy()
    cr, new_r, src_r = ref.kl_track(ref.Orb(), a, b, sel)
    cg, new_g, ok = api.KL_Track(a, b, sel)
    assert cg == cr and cr > 300
    assert {int(s): new_r[j].tobytes() for j, s in enumerate(src_r)} == {int(i): new_g[i].tobytes() for i in np.nonzero(ok)[0]}
