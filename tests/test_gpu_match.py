"""GPU parity: Hamming 2-NN kernels, filters, stereo and projection matching through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from mcvslam_b200 import synth

pytestmark = pytest.mark.gpu


def _eq_dm(a, b, fields=("queryIdx", "trainIdx", "imgIdx", "distance")):
    assert a.shape == b.shape, (a.shape, b.shape)
    for f in fields:
        bad = np.nonzero(a[f] != b[f])
        assert len(bad[0]) == 0, f"{f} differs at {[x[:5] for x in bad]}"


@pytest.mark.parametrize("nq,nt,low", [(2000, 2000, False), (300, 500, True), (1, 2001, False), (129, 1, False), (5, 0, False),
                                       (1000, 70000, True),
                                       # the one-launch warp-per-query kernel (nq >= 256, nt <= 4096, < 2^23 pairs): ragged tiles, ties, nt in {0, 1}
                                       (1024, 257, True), (500, 4096, True), (2047, 4095, False), (4000, 2000, False), (1025, 1, False), (1030, 0, False), (256, 513, True),
                                       (2000, 2000, True)])
def test_knn2_bf(api, oracle, nq, nt, low):
    q = synth.descriptors(nq, 1, low); t = synth.descriptors(nt, 2, low)
    if low:
        q[:, 3:] = 0; t[:, 3:] = 0          # heavy ties
    res = api.Matcher.KnnMatch(q, t)
    ref, k = oracle.knn2_bf(q, t)
    _eq_dm(res.knn, ref[:, :k])
    fp = api.Matcher.KnnMatchRows(q, t)
    _eq_dm(fp.knn, oracle.knn2_firstparty(q, t))
    if nt > 0:
        bm = api.Matcher.BFMatch(q, t)
        _eq_dm(bm.m, ref[:, 0])


def test_knn2_golden(api, golden):
    res = api.Matcher.KnnMatch(golden["g4_q"], golden["g4_t"])
    assert np.array_equal(res.knn["trainIdx"], golden["g4_idx"]) and np.array_equal(res.knn["distance"], golden["g4_dist"])
    res = api.Matcher.KnnMatch(golden["g1_desc0"], golden["g1_desc1"])
    assert np.array_equal(res.knn["trainIdx"], golden["g1_bf_idx"]) and np.array_equal(res.knn["distance"], golden["g1_bf_dist"])


def test_knn2_candidates_and_filters(api, oracle):
    rng = np.random.default_rng(3)
    q = synth.descriptors(700, 5, True); t = synth.descriptors(900, 6, True)
    q[:, 2:] = 0; t[:, 2:] = 0
    lens = rng.integers(0, 70, 700); lens[:5] = [0, 1, 2, 33, 64]
    off = np.zeros(701, np.int32); off[1:] = np.cumsum(lens)
    cidx = rng.integers(0, 900, off[-1]).astype(np.int32)
    res = api.Matcher.KnnMatchCandidates(q, t, off, cidx)
    ref = oracle.knn2_candidates(q, t, off, cidx)
    _eq_dm(res.knn, ref)
    for ratio in (0.6, 0.7, 1.0):
        a = res.FilterRatio(ratio); b = oracle.filter_ratio(ref, ratio)
        _eq_dm(a.m, b)
        _eq_dm(a.FilterThreshold(3).m, oracle.filter_threshold(b, 3))
    # orientation histogram on random angles
    k1 = np.zeros(700, api.KP_DTYPE); k2 = np.zeros(900, api.KP_DTYPE)
    k1["angle"] = rng.uniform(0, 360, 700).astype(np.float32); k2["angle"] = rng.uniform(0, 360, 900).astype(np.float32)
    m = oracle.filter_ratio(oracle.knn2_bf(q, t)[0], 1.0)
    _eq_dm(api.MatchRes(m).FilterOrientation(k1, k2).m, oracle.filter_orientation(m, k1, k2))


def test_dbow_match(api, oracle):
    rng = np.random.default_rng(8)
    d1 = synth.descriptors(400, 1, True); d2 = synth.descriptors(500, 2, True)
    n1 = rng.integers(0, 40, 400); n2 = rng.integers(0, 40, 500)
    fv1 = {}; fv2 = {}
    for i, n in enumerate(n1): fv1.setdefault(int(n) * 3, []).append(i)
    for i, n in enumerate(n2): fv2.setdefault(int(n) * 3 if n % 5 else int(n) * 3 + 1, []).append(i)
    res = api.Matcher.DBowMatch(d1, fv1, d2, fv2).knn
    # restatement of src/Matcher.cpp:146-193 with the oracle's candidate 2-NN
    exp = []
    for nid in sorted(set(fv1) & set(fv2)):
        c = np.array(fv2[nid], np.int32)
        for f1 in fv1[nid]:
            r = oracle.knn2_candidates(d1[f1:f1 + 1], d2, np.array([0, len(c)], np.int32), c)[0]
            if r[1]["distance"] != 999:
                exp.append([(f1, c[r[0]["trainIdx"]], -1, r[0]["distance"]), (f1, c[r[1]["trainIdx"]], -1, r[1]["distance"])])
    exp = np.array(exp, api.DM_DTYPE).reshape(-1, 2)
    _eq_dm(res, exp)


@pytest.mark.parametrize("seed", [21, 22])
def test_stereo(api, oracle, seed):
    trip = synth.triplet(seed)
    EL, ER = api.ORB(2000, 1.2, 8, 28, 15), api.ORB(2000, 1.2, 8, 28, 15)
    OL, OR = oracle.Orb(2000, 1.2, 8, 28, 15), oracle.Orb(2000, 1.2, 8, 28, 15)
    nl, kl, dl = EL.Extract(trip[0]); nr, kr, dr = ER.Extract(trip[1])
    OL.extract(trip[0]); OR.extract(trip[1])
    bf, b = 955.40503, 1.0
    ur, dp, bd, br = api.ComputeStereoMatch(EL, ER, kl, dl, kr, dr, bf, b)
    n, uro, dpo, bdo, bro = oracle.stereo_match(OL, OR, kl, dl, kr, dr, 480, bf, b)
    assert n > 200, "synthetic stereo pair should match"
    assert np.array_equal(bd, bdo) and np.array_equal(br, bro)
    assert np.array_equal(ur.view(np.uint32), uro.view(np.uint32)) and np.array_equal(dp.view(np.uint32), dpo.view(np.uint32))


@pytest.mark.parametrize("shift", [1, 3])
def test_stereo_coarse_row_bins(api, oracle, shift, monkeypatch):
    """The right keypoints' (octave, row) table falls back to coarser row bins when it would not fit the rows kernel's shared memory
    (many levels of a very tall image); the candidate gate stays exact, so the result must not change. MCV_STEREO_ROW_SHIFT forces it."""
    monkeypatch.setenv("MCV_STEREO_ROW_SHIFT", str(shift))
    trip = synth.triplet(23)
    EL, ER = api.ORB(2000, 1.2, 8, 28, 15), api.ORB(2000, 1.2, 8, 28, 15)
    OL, OR = oracle.Orb(2000, 1.2, 8, 28, 15), oracle.Orb(2000, 1.2, 8, 28, 15)
    nl, kl, dl = EL.Extract(trip[0]); nr, kr, dr = ER.Extract(trip[1])
    OL.extract(trip[0]); OR.extract(trip[1])
    bf, b = 955.40503, 1.0
    ur, dp, bd, br = api.ComputeStereoMatch(EL, ER, kl, dl, kr, dr, bf, b)
    n, uro, dpo, bdo, bro = oracle.stereo_match(OL, OR, kl, dl, kr, dr, 480, bf, b)
    assert n > 200
    assert np.array_equal(bd, bdo) and np.array_equal(br, bro)
    assert np.array_equal(ur.view(np.uint32), uro.view(np.uint32)) and np.array_equal(dp.view(np.uint32), dpo.view(np.uint32))


def test_rig_batch(api, oracle):
    frames = np.stack([synth.triplet(s) for s in (31, 32, 33)])
    R = api.Rig()
    out = R.process(frames)
    O = [oracle.Orb(2000, 1.2, 8, 28, 15) for _ in range(3)]
    for f in range(3):
        ks, ds = [], []
        for c in range(3):
            n, k, d = O[c].extract(frames[f, c])
            assert out["counts"][f, c] == n
            assert out["kps"][f, c, :n].tobytes() == k.tobytes() and out["desc"][f, c, :n].tobytes() == d.tobytes()
            ks.append(k); ds.append(d)
        n, ur, dp, bd, br = oracle.stereo_match(O[0], O[1], ks[0], ds[0], ks[1], ds[1], 480, 955.40503, 1.0)
        nl = len(ks[0])
        assert np.array_equal(out["u_right"][f, :nl].view(np.uint32), ur.view(np.uint32))
        assert np.array_equal(out["depth_left"][f, :nl].view(np.uint32), dp.view(np.uint32))
    assert R.last_launches() > 0


def test_projection(api, oracle):
    img = synth.scene(55)
    E = api.ORB(2000, 1.2, 8, 28, 15)
    n, k, d = E.Extract(img)
    rng = np.random.default_rng(9)
    n_mp = 10000
    fx = fy = 955.40503 * 640 / 512; cx, cy = 320.0, 240.0
    th = 0.01
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]], np.float32)
    t = np.array([0.02, -0.01, 0.03], np.float32)
    src = rng.integers(0, n, n_mp)
    z = rng.uniform(2, 50, n_mp).astype(np.float32)
    u = k["x"][src] + rng.normal(0, 2.0, n_mp).astype(np.float32); v = k["y"][src] + rng.normal(0, 2.0, n_mp).astype(np.float32)
    pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1).astype(np.float32)
    pw = ((pc - t) @ R).astype(np.float32)      # Pc = R Pw + t  ->  Pw = R^T (Pc - t)
    pw[:50, 2] = -5 - pw[:50, 2]                 # some behind the camera
    md = d[src].copy()
    flips = rng.integers(0, 41, n_mp)
    for m in range(n_mp):
        bits = rng.choice(256, flips[m], replace=False)
        np.bitwise_xor.at(md[m], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    lvl = k["octave"][src].astype(np.int32)
    for r_th in (5.0, 7.0, 10.0):
        cnt, oi, od = api.ProjectBunchMapPoints(k, d, 640, 480, E.mvScaleFactor, R, t, [fx, fy, cx, cy], pw, md, lvl, r_th)
        cnto, oio, odo = oracle.project_match(k, d, 640, 480, E.mvScaleFactor, R, t, [fx, fy, cx, cy], pw, md, lvl, r_th)
        assert cnt == cnto and cnt > 1000
        assert np.array_equal(oi, oio) and np.array_equal(od, odo)


def _fuse_scene(oracle_or_api_orb, rng, k, d, n_mp):
    """MapPoints placed on / near the keypoints of one extracted image, as Map::Fuse meets them."""
    n = len(k)
    fx = fy = 955.40503 * 640 / 512; cx, cy = 320.0, 240.0
    th = 0.015
    R = np.array([[np.cos(th), 0, np.sin(th)], [0, 1, 0], [-np.sin(th), 0, np.cos(th)]], np.float32)
    t = np.array([0.03, -0.02, 0.05], np.float32)
    Ow = (-R.T @ t).astype(np.float32)
    src = rng.integers(0, n, n_mp)
    z = rng.uniform(2, 50, n_mp).astype(np.float32)
    u = k["x"][src] + rng.normal(0, 1.2, n_mp).astype(np.float32); v = k["y"][src] + rng.normal(0, 1.2, n_mp).astype(np.float32)
    pc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], 1).astype(np.float32)
    pw = ((pc - t) @ R).astype(np.float32)
    pw[:40, 2] = -5 - pw[:40, 2]                                   # behind the camera
    view = pw - Ow
    nrm = (view / np.linalg.norm(view, axis=1, keepdims=True)).astype(np.float32)
    # rotate some normals away: beyond 60 degrees, and some right on the boundary
    ang = rng.uniform(0, np.pi / 2, n_mp); ang[100:140] = np.pi / 3
    perp = np.cross(nrm, np.array([0.3, 1.0, 0.2], np.float32)); perp /= np.linalg.norm(perp, axis=1, keepdims=True)
    nrm = (nrm * np.cos(ang)[:, None] + perp * np.sin(ang)[:, None]).astype(np.float32)
    md = d[src].copy()
    flips = rng.integers(0, 41, n_mp)
    for m in range(n_mp):
        bits = rng.choice(256, flips[m], replace=False)
        np.bitwise_xor.at(md[m], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    lvl = np.clip(k["octave"][src].astype(np.int32) + rng.integers(-1, 2, n_mp), 0, 7).astype(np.int32)
    depth_left = np.where(rng.random(n) < 0.5, rng.uniform(0, 640, n), -1).astype(np.float32)
    # make the "stereo" residual small for a share of them so that branch also passes
    return dict(R=R, t=t, Ow=Ow, K=[fx, fy, cx, cy], pw=pw, nrm=nrm, md=md, lvl=lvl, depth_left=depth_left, u=u, z=z, src=src)


@pytest.mark.gpu
def test_fuse_match(api, oracle):
    """Map::Fuse matching front-end (src/Map.cpp:478-527): CUDA == oracle, every MapPoint."""
    img = synth.scene(56)
    E = api.ORB(2000, 1.2, 8, 28, 15)
    n, k, d = E.Extract(img)
    rng = np.random.default_rng(21)
    S = _fuse_scene(E, rng, k, d, 12000)
    bf = 955.40503
    # stereo branch as written compares ur = u - bf / z with depth_left[idx]: give half of the keypoints a value near it
    ur = S["u"] - bf / S["z"]
    S["depth_left"][S["src"][::2]] = (ur[::2] + rng.normal(0, 0.4, len(ur[::2]))).astype(np.float32)
    s2 = E.mvLevelSigma2; is2 = E.mvInvLevelSigma2
    cnt, oi, od = api.FuseMatch(k, d, 640, 480, s2, is2, S["R"], S["t"], S["Ow"], S["K"], S["depth_left"], bf, S["pw"], S["nrm"], S["md"], S["lvl"])
    cnto, oio, odo = oracle.fuse_match(k, d, 640, 480, s2, is2, S["R"], S["t"], S["Ow"], S["K"], S["depth_left"], bf, S["pw"], S["nrm"], S["md"], S["lvl"])
    assert cnt == cnto and cnt > 500, (cnt, cnto)
    assert np.array_equal(oi, oio) and np.array_equal(od, odo)
    # both gates are exercised: some matched keypoints carry depth, some do not
    m = oi[oi >= 0]
    assert (S["depth_left"][m] >= 0).any() and (S["depth_left"][m] < 0).any()
    # no MapPoints / no keypoints
    assert api.FuseMatch(k[:0], d[:0], 640, 480, s2, is2, S["R"], S["t"], S["Ow"], S["K"], S["depth_left"][:0], bf, S["pw"][:5], S["nrm"][:5], S["md"][:5], S["lvl"][:5])[0] == 0


@pytest.mark.gpu
def test_wnd_track(api, oracle):
    """Tracker::Wnd_Track (src/Tracker.cpp:341-360) between two views of a scene: CUDA == oracle, incl. the reference's
    first-candidate index."""
    a = synth.scene(57)
    b = np.roll(a, (3, -5), (0, 1))                                # second view: small shift
    E = api.ORB(2000, 1.2, 8, 28, 15)
    n1, k1, d1 = E.Extract(a)
    n2, k2, d2 = E.Extract(b)
    rng = np.random.default_rng(4)
    q = np.sort(rng.choice(n1, 1200, replace=False)).astype(np.int32)   # keypoints owning a MapPoint
    cnt, oi, ob, od = api.WndTrack(k1, d1, q, k2, d2, 640, 480)
    cnto, oio, obo, odo = oracle.wnd_track(k1, d1, q, k2, d2, 640, 480)
    assert cnt == cnto and cnt > 100, (cnt, cnto)
    assert np.array_equal(oi, oio) and np.array_equal(ob, obo) and np.array_equal(od, odo)
    assert (oi[oi >= 0] != ob[oi >= 0]).any()                      # the reference's queryIdx slip is visible in the data


@pytest.mark.gpu
def test_bow_transform_on_the_shipped_vocabulary(api):
    """Object::ComputeBow on the reference's own vocabulary (the part of Vocabulary/orbvoc.dbow3 the fixture's descriptors walk
    through) against what the reference's DBoW3 computed: tests/golden/voc_golden.npz, made by tests/golden/make_voc_golden.py."""
    from test_oracle_golden import _voc_fixture
    voc, desc, exp = _voc_fixture()
    g = api.Vocabulary(voc).transform(desc, 4)
    assert np.array_equal(g["bow_ids"], exp["bow_ids"]) and g["bow_vals"].tobytes() == exp["bow_vals"].tobytes()
    assert np.array_equal(g["fv_nodes"], exp["fv_nodes"]) and np.array_equal(g["fv_off"], exp["fv_off"]) and np.array_equal(g["fv_idx"], exp["fv_idx"])


@pytest.mark.gpu
@pytest.mark.parametrize("weighting,norm,K,L,levelsup", [(0, 1, 10, 4, 2), (0, 2, 10, 3, 1), (1, 0, 8, 3, 4), (2, 1, 10, 4, 3), (3, 0, 33, 2, 1)])
def test_bow_transform(api, oracle, weighting, norm, K, L, levelsup):
    """Object::ComputeBow = DBoW3 Vocabulary::transform (modules/DBow3/src/Vocabulary.cpp:572-672): words, nodes, and the two
    vectors (doubles bit for bit) equal the oracle's, on a synthetic vocabulary with sibling ties and stopped words."""
    voc = synth.random_vocabulary(100 + weighting * 7 + K, K=K, L=L, weighting=weighting, norm=norm)
    rng = np.random.default_rng(K + L)
    leaves = np.where(voc["word_id"] >= 0)[0]
    d = voc["node_desc"][rng.choice(leaves, 2000)].copy()           # features near words ...
    flips = rng.integers(0, 60, len(d))
    for m in range(len(d)):
        bits = rng.choice(256, flips[m], replace=False)
        np.bitwise_xor.at(d[m], bits // 8, (1 << (bits % 8)).astype(np.uint8))
    d[-200:] = rng.integers(0, 256, (200, 32), dtype=np.uint8)      # ... and some anywhere
    V = api.Vocabulary(voc)
    a = V.transform(d, levelsup)
    b = oracle.bow_transform(d, voc, levelsup)
    for key in ("word", "nid", "bow_ids", "fv_nodes", "fv_off", "fv_idx"):
        assert np.array_equal(a[key], b[key]), key
    assert a["weight"].tobytes() == b["weight"].tobytes() and a["bow_vals"].tobytes() == b["bow_vals"].tobytes()
    assert len(a["bow_ids"]) > 50 and len(a["fv_nodes"]) >= 1
    # the FeatureVector is what DBowMatch consumes: run it end to end against the oracle's candidate 2-NN
    d2 = d[rng.permutation(len(d))]
    a2 = V.transform(d2, levelsup)
    r = api.Matcher.DBowMatch(d, (a["fv_nodes"], a["fv_off"], a["fv_idx"]), d2, (a2["fv_nodes"], a2["fv_off"], a2["fv_idx"]))
    assert len(r.knn) > 100
    assert V.transform(d[:0], levelsup)["bow_ids"].size == 0
    V.close()


@pytest.mark.parametrize("sizes", [[0, 1, 2, 3, 4, 5, 8, 33, 64, 100, 0, 7], [700], [1] * 50 + [2] * 50, list(range(0, 70))])
def test_distinctive_descriptors(api, oracle, sizes):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cpp:101-150) batched: BestIdx, BestMedian and the cloned row."""
    from test_oracle_golden import _ragged_observations
    desc, off = _ragged_observations(len(sizes), sizes)
    bi, bm, od = api.ComputeDistinctiveDescriptors(desc, off)
    bio, bmo = oracle.distinctive(desc, off)
    assert np.array_equal(bi, bio) and np.array_equal(bm, bmo)
    for m in range(len(sizes)):
        assert np.array_equal(od[m], desc[off[m] + bi[m]] if bi[m] >= 0 else np.zeros(32, np.uint8))


def test_distinctive_descriptors_ties(api, oracle):
    """Identical observations: every median is 0 and the first row wins; an empty batch is a no-op."""
    desc = np.tile(synth.descriptors(1, 5), (9, 1))
    bi, bm, od = api.ComputeDistinctiveDescriptors(desc, [0, 4, 9])
    assert bi.tolist() == [0, 0] and bm.tolist() == [0, 0]
    bi, bm, od = api.ComputeDistinctiveDescriptors(np.zeros((0, 32), np.uint8), [0])
    assert len(bi) == 0


@pytest.mark.parametrize("nq,nt,low", [(1025, 8193, False), (129, 70001, True), (4097, 2049, False), (5000, 5001, True), (1000, 70000, True),
                                       (30000, 300, False), (2049, 4100, True), (8, 1 << 20, False), (300000, 40, True)])
def test_knn2_tensor_core_path(api, oracle, nq, nt, low):
    """Problems of >= 2^23 pairs run on tcgen05 (int8 GEMM of the +-1 expanded descriptors, match_tc_kernels.cu): tile edges
    (M = 128 queries, N = 256 train rows per tile), the split + merge path (few query tiles, long train walk), tie-heavy sets."""
    assert nq * nt >= 1 << 23
    q = synth.descriptors(nq, 21, low); t = synth.descriptors(nt, 22, low)
    if low:
        q[:, 2:] = 0; t[:, 2:] = 0          # 16 significant bits: most nearest neighbours tie, the index order decides
    res = api.Matcher.KnnMatch(q, t)
    if nq * nt <= 400_000_000:
        ref, k = oracle.knn2_bf(q, t)
    else:
        s, ref = oracle.bench_knn2(q, t, 8); k = 2
    _eq_dm(res.knn, ref[:, :k])


def test_knn2_pairs_device(api, oracle):
    """mcv_knn2_pairs_device: BF 2-NN between images of one device-resident descriptor array (configs[3]: consecutive frames),
    ragged per-image counts incl. an empty image and a full one; every pair == the oracle on the two descriptor sets."""
    import torch
    rng = np.random.default_rng(31)
    n_img, cap = 7, 1500
    counts = np.array([1500, 1203, 0, 1, 777, 1499, 300], np.int32)
    desc = rng.integers(0, 256, (n_img, cap, 32), dtype=np.uint8)
    desc[5, :, 2:] = 0; desc[6, :, 2:] = 0                                 # tie-heavy pair
    pq = np.array([0, 1, 2, 3, 4, 5, 6, 0, 5], np.int32); pt = np.array([1, 2, 3, 4, 5, 6, 0, 0, 5], np.int32)
    dev = torch.device("cuda", 0)
    d_desc = torch.from_numpy(desc).to(dev); d_cnt = torch.from_numpy(counts).to(dev)
    d_pq = torch.from_numpy(pq).to(dev); d_pt = torch.from_numpy(pt).to(dev)
    idx = torch.full((len(pq), cap, 2), -7, dtype=torch.int32, device=dev); dst = torch.full((len(pq), cap, 2), -7, dtype=torch.int32, device=dev)
    api._check(api.lib().mcv_knn2_pairs_device(d_desc.data_ptr(), d_cnt.data_ptr(), n_img, cap, d_pq.data_ptr(), d_pt.data_ptr(), len(pq), idx.data_ptr(),
                                               dst.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize()
    idx = idx.cpu().numpy(); dst = dst.cpu().numpy()
    for p in range(len(pq)):
        nq, nt = counts[pq[p]], counts[pt[p]]
        assert (idx[p, nq:] == -7).all()                                    # rows beyond the query image's count are untouched
        if nq == 0:
            continue
        ref, k = oracle.knn2_bf(desc[pq[p], :nq], desc[pt[p], :nt])
        assert np.array_equal(idx[p, :nq, :k], ref["trainIdx"][:, :k]), p
        assert np.array_equal(dst[p, :nq, :k], ref["distance"][:, :k].astype(np.int32)), p
        assert (idx[p, :nq, k:] == -1).all()
