"""CPU: the order-defining host epilogues of the C ABI (mcv_filter_ratio / _threshold / _orientation / _fmatrix — host code inside
libmcv_b200.so, no device needed) against oracle/_ref = the reference's own MatchRes / MatchResKnn methods (src/Matcher.cpp:23-111,
310-325) and against the oracle restatement."""
import numpy as np
import pytest

from test_ref_parity import _matches, ref  # noqa: F401  (fixture)


def _eq(a, b):
    assert a.tobytes() == b.tobytes()


def test_filter_fmatrix_cabi(api, oracle, ref):
    s2 = oracle.Orb().sigma2
    rng = np.random.default_rng(12)
    for seed in range(8):
        k1, k2, m = _matches(oracle, seed=seed + 10)
        if seed < 3:      # pure sideways translation, F = [t]x: matches on (nearly) the same row survive
            F = np.array([[0, 0, 0], [0, 0, -1], [0, 1, 0]], np.float32) * np.float32(rng.uniform(0.5, 2))
            k2 = k2.copy(); sel = rng.random(len(m)) < 0.5
            k2["y"][m["trainIdx"][sel]] = k1["y"][m["queryIdx"][sel]] + rng.normal(0, 1.5, int(sel.sum())).astype(np.float32)
        elif seed == 3:
            F = np.zeros((3, 3), np.float32)                                     # den == 0 rejects everything
        else:
            F = rng.normal(0, 1e-3 if seed % 2 else 1, (3, 3)).astype(np.float32)
        got = api.MatchRes(m).FilterFMatrix(k1, k2, F, s2).m
        _eq(got, ref.filter_fmatrix(m, k1, k2, F, s2))
        _eq(got, oracle.filter_fmatrix(m, k1, k2, F, s2))
        if seed < 3:
            assert 0 < len(got) < len(m)
            assert not np.array_equal(got["queryIdx"], np.sort(got["queryIdx"])) or len(got) < 3   # swap-remove reorders
        if seed == 3:
            assert len(got) == 0
    assert len(api.MatchRes(m[:0]).FilterFMatrix(k1, k2, F, s2).m) == 0
    bad = m.copy(); bad["trainIdx"][0] = len(k2)                                  # out-of-range index: an error code, not a crash
    with pytest.raises(api.McvError):
        api.MatchRes(bad).FilterFMatrix(k1, k2, F, s2)


def test_other_filters_cabi(api, oracle, ref):
    rng = np.random.default_rng(9)
    knn = np.zeros((400, 2), oracle.DM_DTYPE)
    knn["distance"] = rng.integers(0, 120, (400, 2)); knn["distance"][:30] = 0; knn["distance"][30:60, 1] = 999
    knn["queryIdx"] = np.arange(400)[:, None]; knn["trainIdx"] = rng.integers(0, 999, (400, 2))
    for r in (0.6, 0.7, 0.75):
        _eq(api.MatchResKnn(knn).FilterRatio(r).m, ref.filter_ratio(knn, r))
    _eq(api.MatchResKnn(knn[:, :1]).FilterRatio(0.6).m, ref.filter_ratio(knn[:, :1], 0.6))
    for seed in range(4):
        k1, k2, m = _matches(oracle, seed=seed)
        for th in (46, 34, 0, 200):
            _eq(api.MatchRes(m).FilterThreshold(th).m, ref.filter_threshold(m, th))
        if seed == 1:                                     # few distinct rotations: equal-sized bins exercise std::sort's tie order
            k1 = k1.copy(); k2 = k2.copy()
            k2["angle"] = np.float32(10.0); k1["angle"] = (rng.integers(0, 6, len(k1)) * 9 + 10).astype(np.float32)
        _eq(api.MatchRes(m).FilterOrientation(k1, k2).m, ref.filter_orientation(m, k1, k2))
