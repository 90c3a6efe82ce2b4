"""CPU check of the quadtree core shared with the sm_100a kernel (mcvslam_b200/csrc/octree_core.cuh, compiled by g++ through
tests/cpp/octree_core_shim.cpp) against the oracle's DistributeOctTree (ORBextractor.cc:469-580): same survivors, same order."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpp", "octree_core_shim.cpp")
HDR = os.path.join(HERE, "..", "mcvslam_b200", "csrc", "octree_core.cuh")
SO = os.path.join(HERE, "cpp", "_build", "liboctcore.so")


@pytest.fixture(scope="module")
def core():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or max(os.path.getmtime(SRC), os.path.getmtime(HDR)) > os.path.getmtime(SO):
        subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-shared", "-fPIC", "-x", "c++", SRC, "-o", SO], check=True)
    L = C.CDLL(SO)
    L.octcore_distribute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.octcore_distribute_kernel_form.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
    return L


def run_core(L, x, y, r, bw, bh, N):
    pts = (x.astype(np.uint32) | (y.astype(np.uint32) << 12) | (r.astype(np.uint32) << 24)).astype(np.uint32)
    out = np.zeros(N + 8, np.uint32)
    n = L.octcore_distribute(pts.ctypes.data, len(pts), bw, bh, N, out.ctypes.data, len(out))
    assert n >= 0, n
    out = out[:n]
    for order in (0, 1):     # the kernel's form: equal-key tail of the drain by emulated lanes, writes in either lane order
        o2 = np.zeros(N + 8, np.uint32)
        n2 = L.octcore_distribute_kernel_form(pts.ctypes.data, len(pts), bw, bh, N, o2.ctypes.data, len(o2), order)
        assert n2 == n and np.array_equal(o2[:n], out), ("equal-key drain differs", order, n2, n)
    return np.stack([out & 0xfff, (out >> 12) & 0xfff, out >> 24], 1).astype(np.int64)


def run_oracle(O, x, y, r, bw, bh, N):
    from oracle.oracle import KP_DTYPE
    k = np.zeros(len(x), KP_DTYPE)
    k["x"] = x; k["y"] = y; k["response"] = r
    o = O.distribute_octree(k, 16, 16 + bw, 16, 16 + bh, N)
    return np.stack([o["x"], o["y"], o["response"]], 1).astype(np.int64)


def unique_points(rng, n, bw, bh, cluster=False):
    if cluster:
        cx, cy = rng.integers(0, bw, 12), rng.integers(0, bh, 12)
        k = rng.integers(0, 12, 4 * n)
        x = np.clip(cx[k] + rng.integers(-9, 10, 4 * n), 0, bw - 1); y = np.clip(cy[k] + rng.integers(-9, 10, 4 * n), 0, bh - 1)
    else:
        x = rng.integers(0, bw, 4 * n); y = rng.integers(0, bh, 4 * n)
    _, first = np.unique(x.astype(np.int64) * 8192 + y, return_index=True)
    first = np.sort(first)[:n]
    return x[first], y[first]


CASES = [(602, 442, 434, 1468), (602, 442, 434, 300), (495, 362, 362, 1130), (141, 96, 122, 565), (602, 442, 2000, 5000), (1242, 682, 1086, 4612),
         (900, 200, 300, 1200), (480, 480, 2000, 1511), (480, 480, 1500, 1511), (1500, 100, 200, 900), (100, 100, 50, 49), (100, 100, 3, 1), (37, 51, 40, 300), (4000, 3000, 700, 3000)]


@pytest.mark.parametrize("bw,bh,N,M", CASES)
def test_core_matches_oracle(core, oracle, bw, bh, N, M):
    rng = np.random.default_rng(bw * 7 + N)
    for rep in range(6):
        x, y = unique_points(rng, M, bw, bh, cluster=rep >= 3)
        r = rng.integers(15, 40 if rep % 2 else 255, len(x))     # narrow range: many equal responses
        a = run_core(core, x, y, r, bw, bh, N)
        b = run_oracle(oracle, x, y, r, bw, bh, N)
        assert a.shape == b.shape and np.array_equal(a, b), (rep, len(a), len(b))


def test_core_dense_block_and_row_order(core, oracle):
    # every pixel of a block (adjacent points, deepest paths), in row-major (reference candidate) order
    yy, xx = np.mgrid[300:340, 560:602]
    x, y = xx.ravel(), yy.ravel()
    r = np.full(len(x), 20)
    for N in (10, 500, 1680, 4000):
        assert np.array_equal(run_core(core, x, y, r, 602, 442, N), run_oracle(oracle, x, y, r, 602, 442, N))


def test_heap_primitives_match_libstdcxx(core):
    """oct::heap_push / oct::heap_pop against std::push_heap / std::pop_heap on the same
    entries: every heap size from 1 to 70 (all tree shapes, lone-child cases), larger ones, count ranges from all-equal to wide."""
    core.octcore_heap_check.argtypes = [C.c_uint, C.c_int, C.c_int]
    for n in list(range(1, 140)) + [255, 256, 257, 434, 1000, 1511, 2003, 4095, 4096, 4097]:
        for max_count in (1, 2, 3, 8, 1000):
            for seed in range(3):
                assert core.octcore_heap_check(seed + 10 * n, n, max_count) == 0, (n, max_count, seed)
