"""The C++ host mirror (mcvslam_b200/host/mcvslam_b200.hpp: MCVSLAM::ORB / Matcher / Frame / Object over the C ABI) through
its own test program tests/cpp/host_mirror_test.cpp, which compares every result with the CPU oracle bit for bit."""
import os
import subprocess

import pytest


def _bin(oracle):
    import sys
    from mcvslam_b200 import build as B
    B.build()
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "cpp"))
    import build_host_test
    return build_host_test.build()


def test_host_mirror_builds_and_fails_loudly_without_gpu(oracle, tmp_path):
    import torch
    exe = _bin(oracle)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, str(tmp_path), "--no-device"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_host_mirror_equals_oracle(oracle, tmp_path):
    exe = _bin(oracle)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all results equal the oracle" in r.stdout


@pytest.mark.gpu
def test_matcher_entry_points_are_reentrant(oracle, tmp_path):
    """Four host threads inside the matcher entry points at once (per-thread stream + scratch): every result equals the oracle."""
    _bin(oracle)
    import build_host_test
    r = subprocess.run([build_host_test.BIN_THREADS, "4", "12"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all results equal the oracle" in r.stdout
