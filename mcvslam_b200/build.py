"""Builds mcvslam_b200/libmcv_b200.so (the C-ABI engine, include/mcv_b200.h) in-tree with nvcc for sm_100a.

    python -m mcvslam_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU. Objects go to mcvslam_b200/_build/; the .so sits next to this file so that it
travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
SO = os.path.join(HERE, "libmcv_b200.so")
SOURCES = ["engine.cu", "image_kernels.cu", "fast_kernels.cu", "octree_kernels.cu", "describe_kernels.cu", "stereo_kernels.cu", "host_filters.cu",
           "match_kernels.cu", "match_tc_kernels.cu", "lk_kernels.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# --fmad=false: the parity-critical float code uses explicit _rn intrinsics; this is the safety net for everything else.
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--fmad=false", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function"]
# tuning experiments: MCV_NVCC_EXTRA="-DFS_MINB=6" python -m mcvslam_b200.build --force
FLAGS += os.environ.get("MCV_NVCC_EXTRA", "").split()


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "mcv_b200.h"), __file__]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    deps = _deps()
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]

    def compile_one(pair):
        src, obj = pair
        if not force and not _stale(obj, deps):
            return ""
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        logs = list(ex.map(compile_one, zip(SOURCES, objs)))
    if verbose:
        for s, l in zip(SOURCES, logs):
            if l:
                print("==", s, "\n", l)
    if force or _stale(SO, objs):
        cmd = [NVCC, "-shared", "-o", SO] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
