"""Test infrastructure: compiles tests/cpp/host_mirror_test.cpp — the C++ host mirror (mcvslam_b200/host/mcvslam_b200.hpp) driven
like the reference's own test programs — with g++ (no nvcc), linked against the engine and, as the checker, the CPU oracle."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "host_mirror_test.cpp")
BIN = os.path.join(HERE, "_build", "host_mirror_test")
SRC_THREADS = os.path.join(HERE, "matcher_threads_test.cpp")
BIN_THREADS = os.path.join(HERE, "_build", "matcher_threads_test")


def build(force=False):
    _build_one(SRC_THREADS, BIN_THREADS, force)
    return _build_one(SRC, BIN, force)


def _build_one(SRC, BIN, force=False):
    pkg = os.path.join(ROOT, "mcvslam_b200")
    ora = os.path.join(ROOT, "oracle", "_build")
    deps = [SRC, os.path.join(pkg, "host", "mcvslam_b200.hpp"), os.path.join(pkg, "host", "cv_shim.hpp"), os.path.join(ROOT, "include", "mcv_b200.h"),
            os.path.join(pkg, "libmcv_b200.so"), os.path.join(ora, "liborb_oracle.so")]
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    if force or not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", "-Wall", "-pthread", SRC, "-o", BIN, "-L" + pkg, "-lmcv_b200", "-L" + ora, "-lorb_oracle",
               "-Wl,-rpath,$ORIGIN/../../../mcvslam_b200", "-Wl,-rpath,$ORIGIN/../../../oracle/_build"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host mirror test failed to compile:\n%s\n%s" % (r.stdout, r.stderr))
    return BIN
