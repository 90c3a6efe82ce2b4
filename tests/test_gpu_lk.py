"""GPU parity: KL_Track's optical flow (src/Frame.cpp:34-76) through the C ABI vs the CPU oracle and the cv2 fixture."""
import numpy as np
import pytest

from mcvslam_b200 import synth
from test_oracle_golden import _lk_cases

pytestmark = pytest.mark.gpu


def _same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def _smooth_scene(oracle):
    return oracle.gauss7(oracle.gauss7(synth.scene(58)))


def test_lk_track_vs_oracle_and_cv2(api, oracle):
    """Every fixture pair: status, positions and residuals bit-exact against the oracle on ALL points, and against the real
    cv2 where OpenCV's own reads are defined (see test_lk_oracle_against_cv2_fixture)."""
    for k, a, b, pts, nxt, st, err in _lk_cases():
        o, so, eo = api.LkTrack(a, b, pts)
        ro, rs, re = oracle.lk_track(a, b, pts)
        assert np.array_equal(so, rs), k
        assert _same(o, ro) and _same(eo, re), k
        defined = np.floor(pts[:, 1] - np.float32(4.5)) < a.shape[0] - 1
        assert np.array_equal(so[defined], st[defined])
        ok = defined & (st == 1)
        assert _same(o[ok], nxt[ok]) and _same(eo[ok], err[ok]), k


@pytest.mark.parametrize("shape", [(480, 640), (720, 1280), (37, 45), (21, 400)])
def test_lk_track_shapes(api, oracle, shape):
    """Other geometries incl. one whose second level is not larger than the window (maxLevel falls back to 0) and strided input."""
    h, w = shape
    big = synth.scene(77, w + 8, h)
    a = big[:, :w]                                   # stride w + 8
    b = np.ascontiguousarray(synth.shifted(np.ascontiguousarray(a), 0.8, -1.3, 5))
    rng = np.random.default_rng(3)
    pts = np.stack([rng.uniform(-12, w + 12, 700), rng.uniform(-12, h + 12, 700)], 1).astype(np.float32)
    o, so, eo = api.LkTrack(np.ascontiguousarray(a), b, pts)
    ro, rs, re = oracle.lk_track(np.ascontiguousarray(a), b, pts)
    assert np.array_equal(so, rs) and _same(o, ro) and _same(eo, re)
    o0, s0, e0 = api.LkTrack(np.ascontiguousarray(a), b, np.zeros((0, 2), np.float32))
    assert len(o0) == 0


def test_kl_track(api, oracle):
    """KL_Track on extracted keypoints: ok flags, appended keypoints (pt = next, octave 0, other fields kept), the < 10 gate."""
    a = _smooth_scene(oracle); b = synth.shifted(a, 2.4, 1.1, 9, noise=0)   # err < 1 needs a mean residual below one grey level
    E = api.ORB(2000, 1.2, 8, 28, 15)
    n, k, d = E.Extract(a)
    sel = k[::3].copy()
    cnt, new, ok = api.KL_Track(a, b, sel)
    cnto, newo, oko, _, _, _ = oracle.kl_track(a, b, sel)
    assert cnt == cnto and cnt > len(sel) // 2
    assert np.array_equal(ok, oko) and new[ok == 1].tobytes() == newo[oko == 1].tobytes()
    assert (new["octave"][ok == 1] == 0).all() and np.array_equal(new["angle"][ok == 1], sel["angle"][ok == 1])
    cnt, new, ok = api.KL_Track(a, b, sel[:9])
    assert cnt == 0 and not ok.any()


def test_lk_track_batch(api, oracle):
    """mcv_lk_track_batch: the four 640x480 fixture pairs (ragged point lists, one of them empty) in one call == one oracle call per pair."""
    cases = [c for c in _lk_cases() if c[1].shape == (480, 640)]
    prev = np.stack([c[1] for c in cases]); nxt = np.stack([c[2] for c in cases])
    pts = [c[3][:1500 - 211 * k] for k, c in enumerate(cases)]
    pts[2] = pts[2][:0]
    o, st, err = api.LkTrackBatch(prev, nxt, pts)
    for k, c in enumerate(cases):
        ro, rs, re = oracle.lk_track(c[1], c[2], pts[k])
        assert np.array_equal(st[k], rs) and _same(o[k], ro) and _same(err[k], re), k
