"""GPU parity on the edge cases and at BASELINE.json's full sizes: degenerate and crowded
images, odd geometries, the 1280x720 / 5000-feature batch with consecutive-frame matching, the device-resident async rig
path, and the 1M x 1M match through size-independent properties plus sampled oracle rows."""
import numpy as np
import pytest
import torch

from mcvslam_b200 import shard, synth

pytestmark = pytest.mark.gpu


def _same(k, d, ko, do, tag):
    assert len(k) == len(ko), f"{tag}: count {len(k)} vs {len(ko)}"
    assert k.tobytes() == ko.tobytes(), f"{tag}: keypoints differ"
    assert d.tobytes() == do.tobytes(), f"{tag}: descriptors differ"


def test_flat_and_crowded_images(api, oracle):
    rng = np.random.default_rng(3)
    cases = {
        "flat": (np.full((240, 320), 77, np.uint8), (500, 1.2, 4, 20, 7)),                      # no corner anywhere -> 0 keypoints
        "noise": (rng.integers(0, 256, (240, 320), dtype=np.uint8), (1500, 1.2, 4, 9, 2)),      # corners everywhere: crowded cells,
        "salt": ((rng.random((300, 400)) < 0.02).astype(np.uint8) * 255, (3000, 1.2, 3, 20, 7)),  # list / arena capacities
        "checker": (((np.add.outer(np.arange(240) // 5, np.arange(320) // 5) % 2) * 200 + 20).astype(np.uint8), (800, 1.3, 3, 28, 15)),
    }
    # thresholds >= 128 take the other instantiation of k_fast_score (the per-byte "> T" test flips from OR to AND of its carry terms)
    cases["checker_hi"] = (cases["checker"][0], (800, 1.3, 3, 180, 130))
    cases["salt_hi"] = (cases["salt"][0], (3000, 1.2, 3, 150, 60))
    cases["noise_hi"] = (cases["noise"][0], (1500, 1.2, 4, 128, 127))
    for name, (img, prm) in cases.items():
        E = api.ORB(*prm); O = oracle.Orb(*prm, debug=True)
        n, k, d = E.Extract(img)
        no, ko, do = O.extract(img)
        for l in range(prm[2]):
            a = E.debug_level_keypoints(l, 0); b = O.debug_kps(0, l)
            assert len(a) == len(b) and (a["x"] == b["x"]).all() and (a["y"] == b["y"]).all() and (a["response"] == b["response"]).all(), f"{name}: candidates level {l}"
        _same(k, d, ko, do, name)
    assert api.ORB(500, 1.2, 4, 20, 7).Extract(cases["flat"][0])[0] == 0


@pytest.mark.parametrize("w,h,prm", [(641, 479, (1200, 1.2, 8, 28, 15)), (1000, 333, (1500, 1.2, 5, 20, 7)), (250, 200, (300, 1.1, 3, 20, 7)),
                                      (768, 576, (2000, 1.5, 4, 28, 15)), (512, 512, (2000, 1.2, 1, 28, 15)),
                                      (330, 250, (500, 1.2, 8, 20, 7))])   # last: level pitches below the 160-byte NMS / 64-byte rBRIEF TMA boxes
def test_odd_geometries(api, oracle, w, h, prm):
    img = synth.scene(w * 7 + h, w, h)
    E = api.ORB(*prm); O = oracle.Orb(*prm)
    n, k, d = E.Extract(img)
    no, ko, do = O.extract(img)
    assert n > 100
    _same(k, d, ko, do, f"{w}x{h}")
    for l in range(prm[2]):
        assert np.array_equal(E.mvImagePyramid(l), O.level(l))


def test_c4_batch_1280x720_consecutive_matches(api, oracle):
    """configs[3] at full image size / feature count on a small batch: extract + BF 2-NN between consecutive frames."""
    n_frames = 4
    imgs = np.stack([synth.scene(s, 1280, 720) for s in range(n_frames)])
    E = api.ORB(5000, 1.2, 8, 28, 15); O = oracle.Orb(5000, 1.2, 8, 28, 15)
    res = E.ExtractBatch(imgs)
    ref = []
    for i in range(n_frames):
        no, ko, do = O.extract(imgs[i])
        _same(res[i][0], res[i][1], ko, do, f"frame {i}")
        ref.append(do)
        assert no >= 5000
    pb, pe, fb, fe = shard.consecutive_pairs(n_frames, 0, 1)
    assert (pb, pe, fb, fe) == (0, n_frames - 1, 0, n_frames)
    for i in range(pb, pe):
        r = api.Matcher.KnnMatch(res[i][1], res[i + 1][1]).knn
        ro, k = oracle.knn2_bf(ref[i], ref[i + 1])
        assert np.array_equal(r["trainIdx"], ro["trainIdx"]) and np.array_equal(r["distance"], ro["distance"])
        good = api.MatchResKnn(r).FilterRatio(0.6).FilterThreshold(46)
        assert good.m.tobytes() == oracle.filter_threshold(oracle.filter_ratio(ro, 0.6), 46).tobytes()


def test_async_rig_equals_sync(api):
    frames = np.stack([synth.triplet(40 + s) for s in range(5)])
    R = api.Rig()
    ref = R.process(frames)
    dev = torch.device("cuda", 0)
    cap = R.cap
    d_img = torch.from_numpy(frames).to(dev)
    d_k = torch.zeros(5 * 3 * cap * 28, dtype=torch.uint8, device=dev); d_d = torch.zeros(5 * 3 * cap * 32, dtype=torch.uint8, device=dev)
    d_c = torch.zeros(15, dtype=torch.int32, device=dev); d_u = torch.zeros(5 * cap, dtype=torch.float32, device=dev); d_z = torch.zeros(5 * cap, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    for _ in range(2):      # second call exercises the slot rotation
        R.process_async(d_img.data_ptr(), 5, 640, 480, d_k.data_ptr(), d_d.data_ptr(), d_c.data_ptr(), d_u.data_ptr(), d_z.data_ptr())
    R.join(); R.sync()
    cnt = d_c.cpu().numpy().reshape(5, 3)
    assert np.array_equal(cnt, ref["counts"])
    k = d_k.cpu().numpy().view(api.KP_DTYPE).reshape(5, 3, cap); d = d_d.cpu().numpy().reshape(5, 3, cap, 32)
    u = d_u.cpu().numpy().reshape(5, cap)
    for f in range(5):
        for c in range(3):
            n = cnt[f, c]
            assert k[f, c, :n].tobytes() == ref["kps"][f, c, :n].tobytes() and d[f, c, :n].tobytes() == ref["desc"][f, c, :n].tobytes()
        n = cnt[f, 0]
        assert np.array_equal(u[f, :n].view(np.uint32), ref["u_right"][f, :n].view(np.uint32))


def test_sharded_knn2_tiles_equal_oracle(api, oracle):
    q = synth.descriptors(4000, 11, True); t = synth.descriptors(50000, 12, True)
    q[:, 4:] = 0; t[:, 4:] = 0
    dev = torch.device("cuda", 0)
    idx, dst = shard.knn2_sharded(torch.from_numpy(q).to(dev), torch.from_numpy(t).to(dev), shard.engine_knn2_fn(), tile=1 << 14)
    ref, k = oracle.knn2_bf(q, t)
    assert np.array_equal(idx.cpu().numpy(), ref["trainIdx"]) and np.array_equal(dst.cpu().numpy(), ref["distance"].astype(np.int32))


def test_c5_one_million_by_one_million(api, oracle):
    """configs[4] at full size on one GPU: 2^20 x 2^20 descriptors (1.1e12 pairs). Properties: matching a set against itself
    returns every row as its own nearest neighbour at distance 0, the runner-up is never closer, and 48 sampled query rows
    equal the oracle's full scan (distance and index, lexicographic ties)."""
    n = 1 << 20
    g = torch.Generator(device="cpu"); g.manual_seed(5)
    t_cpu = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g)
    dev = torch.device("cuda", 0)
    t = t_cpu.to(dev)
    idx, dst = shard.knn2_sharded(t, t, shard.engine_knn2_fn(), gather=False)
    torch.cuda.synchronize()
    idx = idx.cpu().numpy(); dst = dst.cpu().numpy()
    # random 256-bit rows are distinct with overwhelming probability; duplicates (if any) must still resolve to the lower index
    assert (dst[:, 0] == 0).all()
    self_hit = idx[:, 0] == np.arange(n)
    assert self_hit.mean() > 0.999999 and (idx[~self_hit, 0] < np.arange(n)[~self_hit]).all()
    assert (dst[:, 1] >= dst[:, 0]).all() and (idx[:, 1] != idx[:, 0]).all()
    rows = np.random.default_rng(1).integers(0, n, 48)
    tn = t_cpu.numpy()
    ref, k = oracle.knn2_bf(tn[rows], tn)
    assert np.array_equal(idx[rows], ref["trainIdx"]) and np.array_equal(dst[rows], ref["distance"].astype(np.int32))


@pytest.mark.gpu
def test_rig_submit_wait_matches_sync_call(api):
    """mcv_rig_submit / mcv_rig_wait (asynchronous host-buffer steps, several in flight) returns what the synchronous
    mcv_rig_process returns for the same triplets."""
    import torch
    from mcvslam_b200 import synth
    frames = np.stack([synth.triplet(300 + s) for s in range(6)])
    rig = api.Rig(device=0)
    ref = rig.process(frames)
    cap = rig.cap
    n = len(frames)
    h_img = torch.from_numpy(frames).pin_memory()
    outs, tickets = [], []
    for k in range(5):   # more submits than slots, each on its own result buffers, all in flight at once
        o = dict(kps=torch.zeros(n * 3 * cap * 28, dtype=torch.uint8).pin_memory(), desc=torch.zeros(n * 3 * cap * 32, dtype=torch.uint8).pin_memory(),
                 cnt=torch.zeros(n * 3, dtype=torch.int32).pin_memory(), ur=torch.zeros(n * cap, dtype=torch.float32).pin_memory(),
                 dp=torch.zeros(n * cap, dtype=torch.float32).pin_memory())
        outs.append(o)
        tickets.append(rig.submit(h_img.data_ptr(), n, 640, 480, o["kps"].data_ptr(), o["desc"].data_ptr(), o["cnt"].data_ptr(), o["ur"].data_ptr(),
                                  o["dp"].data_ptr()))
    for k in reversed(range(5)):
        rig.wait(tickets[k])
        o = outs[k]
        cnt = o["cnt"].numpy().reshape(n, 3)
        assert np.array_equal(cnt, ref["counts"])
        kps = o["kps"].numpy().reshape(n, 3, cap, 28); desc = o["desc"].numpy().reshape(n, 3, cap, 32)
        for f in range(n):
            for c in range(3):
                m = cnt[f, c]
                assert kps[f, c, :m].tobytes() == ref["kps"][f, c, :m].tobytes()
                assert desc[f, c, :m].tobytes() == ref["desc"][f, c, :m].tobytes()
            m = cnt[f, 0]
            assert np.array_equal(o["ur"].numpy().reshape(n, cap)[f, :m].view(np.uint32), ref["u_right"][f, :m].view(np.uint32))
    rig.close()


@pytest.mark.gpu
def test_rig_small_chunks_replay_captured_graph(api, oracle):
    """Chunks of <= 32 frames replay a CUDA graph captured on the second call with the same shape (engine.cu: rig_chunk). Five
    calls on ONE rig with different images each time — direct, capture, three replays — must all equal the oracle; then the
    same for the one-triplet call (the reference's Frame-per-call pattern) and after the workspace re-grows in between."""
    rig = api.Rig()
    orbs = [oracle.Orb() for _ in range(3)]

    def check(frames, out):
        for f in range(len(frames)):
            ks, ds = [], []
            for c in range(3):
                n, k, d = orbs[c].extract(frames[f, c])
                assert out["counts"][f, c] == n
                assert out["kps"][f, c, :n].tobytes() == k.tobytes() and out["desc"][f, c, :n].tobytes() == d.tobytes()
                assert not out["kps"][f, c, n:].view(np.uint8).any() and not out["desc"][f, c, n:].any()      # slots behind the count are zero
                ks.append(k); ds.append(d)
            n, ur, dp, bd, br = oracle.stereo_match(orbs[0], orbs[1], ks[0], ds[0], ks[1], ds[1], 480, 955.40503, 1.0)
            assert out["u_right"][f, :len(ks[0])].tobytes() == ur.tobytes()
            assert (out["u_right"][f, len(ks[0]):] == -1).all() and (out["depth_left"][f, len(ks[0]):] == -1).all()

    for call in range(5):
        frames = np.stack([synth.triplet(300 + 2 * call), synth.triplet(301 + 2 * call)])
        check(frames, rig.process(frames))
    for call in range(4):
        frames = np.stack([synth.triplet(320 + call)])
        check(frames, rig.process(frames))
    big = np.stack([synth.triplet(330 + i) for i in range(7)])      # 3 chunks of 3 / 3 / 1 frames: grows the slots' workspaces
    check(big[:2], {k: v[:2] for k, v in rig.process(big).items()})
    for call in range(3):
        frames = np.stack([synth.triplet(340 + call)])
        check(frames, rig.process(frames))


@pytest.mark.gpu
def test_extract_batch_async_device_resident(api, oracle):
    """mcv_orb_extract_batch_async (device in / out, enqueue only on the handle's stream) == the oracle per image; two handles
    on two streams running concurrently do not disturb each other."""
    dev = torch.device("cuda", 0)
    L = api.lib()
    imgs = [np.stack([synth.scene(400 + 10 * k + i) for i in range(3)]) for k in range(2)]
    lanes = []
    for k in range(2):
        s = torch.cuda.Stream(device=dev)
        E = api.ORB(2000, 1.2, 8, 28, 15, device=0, stream=s.cuda_stream)
        cap = E.max_keypoints(0, 640, 480)
        d_img = torch.from_numpy(imgs[k]).to(dev)
        d_k = torch.zeros(3 * cap * 28, dtype=torch.uint8, device=dev); d_d = torch.zeros((3, cap, 32), dtype=torch.uint8, device=dev)
        d_c = torch.zeros(3, dtype=torch.int32, device=dev)
        lanes.append((s, E, cap, d_img, d_k, d_d, d_c))
    torch.cuda.synchronize()
    for rep in range(2):
        for s, E, cap, d_img, d_k, d_d, d_c in lanes:
            api._check(L.mcv_orb_extract_batch_async(E._h, d_img.data_ptr(), 3, 640, 480, d_k.data_ptr(), d_d.data_ptr(), d_c.data_ptr(), cap))
    torch.cuda.synchronize()
    O = oracle.Orb(2000, 1.2, 8, 28, 15)
    for k, (s, E, cap, d_img, d_k, d_d, d_c) in enumerate(lanes):
        cnt = d_c.cpu().numpy(); kps = d_k.cpu().numpy().view(api.KP_DTYPE).reshape(3, cap); desc = d_d.cpu().numpy()
        for i in range(3):
            n, ko, do = O.extract(imgs[k][i])
            assert cnt[i] == n and kps[i, :n].tobytes() == ko.tobytes() and desc[i, :n].tobytes() == do.tobytes(), (k, i)
