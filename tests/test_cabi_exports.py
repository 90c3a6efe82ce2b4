"""CPU: the C-ABI library builds, loads, exports every symbol include/mcv_b200.h declares, and fails loudly (no CPU
fallback) when there is no CUDA device. No compute is attempted here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "mcv_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mcv_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(api):
    L = api.lib()
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mcv_b200.h but not exported"
    assert set(api.EXPORTS) == set(names)


def test_pod_layouts(api):
    assert api.KP_DTYPE.itemsize == 28 and api.DM_DTYPE.itemsize == 16     # cv::KeyPoint / cv::DMatch
    assert ctypes.sizeof(api.OrbParams) == 20 and ctypes.sizeof(api.RigParams) == 28


def test_no_cpu_fallback(api):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert api.lib().mcv_device_count() == 0
    with pytest.raises(api.McvError) as e:
        api.ORB()
    assert e.value.status == -6          # MCV_ERR_NO_DEVICE
    with pytest.raises(api.McvError):
        api.Matcher.KnnMatch(np.zeros((4, 32), np.uint8), np.zeros((4, 32), np.uint8))


def test_host_filters_match_oracle(api, oracle):
    """The order-defining filter epilogues are host code inside the library; they run without a GPU."""
    rng = np.random.default_rng(1)
    knn = np.zeros((500, 2), api.DM_DTYPE)
    knn["distance"] = np.sort(rng.integers(0, 80, (500, 2)), axis=1); knn["queryIdx"] = np.arange(500)[:, None]
    knn["trainIdx"] = rng.integers(0, 400, (500, 2)); knn["distance"][:20] = 0
    for ratio in (0.6, 0.7):
        a = api.MatchResKnn(knn).FilterRatio(ratio); b = oracle.filter_ratio(knn, ratio)
        assert a.m.tobytes() == b.tobytes()
        assert a.FilterThreshold(34).m.tobytes() == oracle.filter_threshold(b, 34).tobytes()
    k1 = np.zeros(500, api.KP_DTYPE); k2 = np.zeros(400, api.KP_DTYPE)
    k1["angle"] = rng.uniform(0, 360, 500).astype(np.float32); k2["angle"] = (rng.integers(0, 12, 400) * 30).astype(np.float32)
    m = oracle.filter_ratio(knn, 1.0)
    assert api.MatchRes(m).FilterOrientation(k1, k2).m.tobytes() == oracle.filter_orientation(m, k1, k2).tobytes()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mcvslam_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "ora_primitives" not in src and "liborb_oracle" not in src, f
