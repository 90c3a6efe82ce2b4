"""Builds oracle/_ref/libmcv_ref.so: the REFERENCE'S OWN hot-path translation units, compiled UNMODIFIED from where they lie
under /root/reference, against oracle/ref_shim (a mini OpenCV / boost / Eigen / pyp / OSG stand-in — none of those libraries
exist in this image; the OpenCV image primitives route to the cv2-pinned models of ora_primitives.hpp / ora_lk.hpp).

    python oracle/build_ref.py [--force]

ORACLE — TEST INFRASTRUCTURE ONLY. Outputs go to oracle/_ref/ only (git-ignored, NOT gpurun-ignored: the .so travels to the GPU
box, where /root/reference does not exist and build() just reuses the prebuilt file). No reference source is copied into the
repo. Flags follow the reference (CMakeLists.txt:4-6,18-19): -O3, no -march, no -ffast-math, libstdc++; -std=c++17 instead of
the reference's C++11 because the boost::shared_mutex stand-in is std::shared_mutex. `-include cstdint` only supplies the
<cstdint> that DBoW3's BowVector/FeatureVector/DescManip get transitively from the real OpenCV headers.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MCV_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SO = os.path.join(OUT, "libmcv_ref.so")
SHIM = os.path.join(HERE, "ref_shim")

# (path under the reference root, what it contributes to the path)
REF_SOURCES = [
    "modules/local_feature/ORB/orb3_extractor/ORBextractor.cc",   # A0-A8: pyramid, per-cell FAST, quadtree, IC_Angle, rBRIEF, operator()
    "modules/local_feature/ORB/ORBExtractor.cpp",                 # A9: MCVSLAM::ORB wrapper + Parse
    "src/Matcher.cpp",                                            # A10-A12: KnnMatch family, DBowMatch, the four filters
    "src/Frame.cpp",                                              # A13: Frame ctor (ThreadPool(3) + grid + ComputeStereoMatch), KL_Track
    "src/Object.cpp",                                             # A14: grid, GetFeaturesInArea, ProjectBunchMapPoints, ComputeBow
    "src/MapPoint.cpp",                                           # f4: ComputeDistinctiveDescriptors
    "src/Map.cpp",                                                # f1: Map::Fuse, Map::ComputeF12
    "src/Tracker.cpp",                                            # f1: Wnd_Track, Bow_Track
    "modules/camera/Pinhole.cpp",
    "modules/thread_pool/thread_pool.cpp",
    "modules/DBow3/src/Vocabulary.cpp",                           # f2: DBoW3::Vocabulary::transform + binary loader
    "modules/DBow3/src/BowVector.cpp",
    "modules/DBow3/src/FeatureVector.cpp",
    "modules/DBow3/src/ScoringObject.cpp",
    "modules/DBow3/src/DescManip.cpp",
    "modules/DBow3/src/quicklz.c",
]
INCLUDES = [SHIM] + [os.path.join(REF, p) for p in ("include", "modules/local_feature/ORB", "modules/local_feature", "modules/local_feature/BaseExtractor",
                                                    "modules/camera", "modules/DBow3/src", "modules/thread_pool")]
CXXFLAGS = ["-O3", "-fPIC", "-std=c++17", "-pthread", "-w", "-include", "cstdint"]
CFLAGS = ["-O3", "-fPIC", "-w"]


def available():
    return os.path.exists(os.path.join(REF, REF_SOURCES[0]))


def _shim_files():
    out = []
    for d, _, fs in os.walk(SHIM):
        out += [os.path.join(d, f) for f in fs]
    return out + [os.path.join(HERE, "ora_primitives.hpp"), os.path.join(HERE, "ora_lk.hpp"), __file__]


def build(force=False):
    """Returns the path of the .so, or None when neither the reference sources nor a prebuilt .so exist."""
    if not available():
        return SO if os.path.exists(SO) else None
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    shim = _shim_files()
    inc = [a for p in INCLUDES for a in ("-I", p)]

    def one(src):
        path = os.path.join(REF, src) if not os.path.isabs(src) else src
        obj = os.path.join(OUT, "obj", os.path.basename(src) + ".o")
        deps = [path] + shim
        if not force and os.path.exists(obj) and all(os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps):
            return obj
        cmd = (["gcc"] + CFLAGS if src.endswith(".c") else ["g++"] + CXXFLAGS) + inc + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("compiling %s failed:\n%s" % (src, r.stderr[-4000:]))
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, REF_SOURCES + [os.path.join(SHIM, "ref_capi.cpp")]))
    if force or not os.path.exists(SO) or any(os.path.getmtime(o) > os.path.getmtime(SO) for o in objs):
        r = subprocess.run(["g++", "-shared", "-pthread", "-o", SO] + objs + ["-Wl,--no-undefined"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("linking libmcv_ref.so failed:\n%s" % r.stderr[-4000:])
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
