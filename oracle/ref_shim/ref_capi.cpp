// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under mcvslam_b200/ may include, link or call this.
//
// C interface (ctypes: oracle/ref.py) over the REFERENCE'S OWN classes, compiled unmodified from /root/reference into
// oracle/_ref/libmcv_ref.so by oracle/build_ref.py: ORB_SLAM3::ORBextractor / MCVSLAM::ORB, MCVSLAM::Matcher + MatchRes /
// MatchResKnn filters, MCVSLAM::Frame (the real constructor: ThreadPool(3) extraction + AssignFeaturesToGrid +
// ComputeStereoMatch), MCVSLAM::Object (grid, ProjectBunchMapPoints, ComputeBow), MCVSLAM::MapPoint
// (ComputeDistinctiveDescriptors), KL_Track, Map::Fuse / Map::ComputeF12, Tracker::Wnd_Track / Bow_Track, DBoW3::Vocabulary.
// This file only marshals flat arrays into those classes and back; it contains no algorithm of the path.
#include <cstdint>
#include <cstring>
#include <chrono>
#include <fstream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "Frame.hpp"
#include "Map.hpp"
#include "MapPoint.hpp"
#include "Matcher.hpp"
#include "ORBExtractor.hpp"
#include "Object.hpp"
#include "Pinhole.hpp"
#include "Tracker.hpp"
#include "pyp/timer/timer.hpp"

using namespace MCVSLAM;

namespace MCVSLAM {
uint KL_Track(ObjectRef obj1, ObjectRef obj2, std::unordered_map<MapPointRef, uint>& mp2idx);  // src/Frame.cpp:34 (external linkage, no header)
}

namespace {
const float kUnitZ[3] = {0.f, 0.f, 1.f};
struct OrbPeek : public ORB {  // protected members of ORB_SLAM3::ORBextractor (ORBextractor.h:76-91)
    using ORB_SLAM3::ORBextractor::mnFeaturesPerLevel;
    using ORB_SLAM3::ORBextractor::umax;
};
cv::Mat wrap_u8(const uint8_t* p, int rows, int cols, size_t step) { return cv::Mat(rows, cols, CV_8UC1, (void*)p, step); }
std::vector<cv::Mat> rows_of(const cv::Mat& m) { std::vector<cv::Mat> v; for (int i = 0; i < m.rows; ++i) v.push_back(m.row(i)); return v; }
int put(const std::vector<cv::DMatch>& m, cv::DMatch* out, int cap) { int n = (int)m.size(); for (int i = 0; i < n && i < cap; ++i) out[i] = m[i]; return n; }
Pinhole* make_cam(const float* intr) { return new Pinhole(std::vector<float>{intr[0], intr[1], intr[2], intr[3]}, cv::Mat::zeros(1, 5, CV_32F)); }
cv::Mat pose_from(const float* Rcw, const float* tcw) {
    cv::Mat T = cv::Mat::eye(4, 4, CV_32F);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T.at<float>(i, j) = Rcw[3 * i + j]; T.at<float>(i, 3) = tcw[i]; }
    return T;
}
struct ObjBox {  // an Object with caller keypoints / descriptors
    std::unique_ptr<Pinhole> cam;
    ObjectRef obj;
};
ObjBox* make_obj(ORB* ext, const cv::KeyPoint* kps, const uint8_t* desc, int n, int w, int h, const float* intr, const float* Rcw, const float* tcw) {
    ObjBox* b = new ObjBox;
    const float dflt[4] = {500.f, 500.f, w / 2.f, h / 2.f};
    b->cam.reset(make_cam(intr ? intr : dflt));
    cv::Mat img(h, w, CV_8UC1, cv::Scalar(0));
    b->obj = std::make_shared<Object>(b->cam.get(), img, ext, CAM_NAME::L);
    b->obj->kps.assign(kps, kps + n);
    b->obj->desps = wrap_u8(desc, n, 32, 32).clone();
    if (Rcw) b->obj->SetPose(pose_from(Rcw, tcw));
    b->obj->AssignFeaturesToGrid();
    return b;
}
MapPointRef make_mp(const float* xyz, const uint8_t* desc, int level, uint id) {
    return std::make_shared<MapPoint>(xyz[0], xyz[1], xyz[2], wrap_u8(desc, 1, 32, 32).clone(), (uint)level, 0u, id, CAM_NAME::L, MP_TYPE::STEREO, 7u);
}
}  // namespace

extern "C" {

const char* ref_version() { return "Sologala/MCVSLAM reference sources compiled against oracle/ref_shim (mini_cv)"; }
void ref_set_num_threads(int n) { cv::setNumThreads(n); }

// ---- extractor: MCVSLAM::ORB(config_path) -> Parse + init (ORBExtractor.cpp:10-23) -------------------------------------
void* ref_orb_create(const char* yaml_path) { try { return new ORB(std::string(yaml_path)); } catch (std::exception& e) { fprintf(stderr, "ref_orb_create: %s\n", e.what()); return nullptr; } }
void ref_orb_destroy(void* h) { delete (ORB*)h; }
int ref_orb_levels(void* h) { return ((ORB*)h)->GetLevels(); }
void ref_orb_params(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota, int* umax16) {
    OrbPeek* E = (OrbPeek*)h;
    const int n = E->GetLevels();
    for (int i = 0; i < n; ++i) { scale[i] = E->mvScaleFactor[i]; inv_scale[i] = E->mvInvScaleFactor[i]; sigma2[i] = E->mvLevelSigma2[i]; inv_sigma2[i] = E->mvInvLevelSigma2[i]; quota[i] = E->mnFeaturesPerLevel[i]; }
    for (int i = 0; i < 16; ++i) umax16[i] = E->umax[i];
}
// BaseExtractor::Extract(img, kps, desps): kps in/out (n_seeds valid on entry). Returns the count (or -1 / -2 = capacity).
int ref_orb_extract(void* h, const uint8_t* img, int w, int hgt, int stride, cv::KeyPoint* kps, int n_seeds, uint8_t* desc, int cap) {
    ORB* E = (ORB*)h;
    std::vector<cv::KeyPoint> v(kps, kps + n_seeds);
    cv::Mat d;
    cv::Mat im = (w > 0 && hgt > 0) ? wrap_u8(img, hgt, w, (size_t)stride) : cv::Mat();
    int n = E->Extract(im, v, d);
    if (n < 0) return n;
    if (n > cap) return -2;
    for (int i = 0; i < n; ++i) kps[i] = v[i];
    for (int i = 0; i < n; ++i) memcpy(desc + 32 * (size_t)i, d.ptr<uchar>(i), 32);
    return n;
}
int ref_orb_level_size(void* h, int level, int* w, int* hgt) { ORB* E = (ORB*)h; *w = E->mvImagePyramid[level].cols; *hgt = E->mvImagePyramid[level].rows; return 0; }
void ref_orb_level_copy(void* h, int level, uint8_t* dst) {
    const cv::Mat& m = ((ORB*)h)->mvImagePyramid[level];
    for (int y = 0; y < m.rows; ++y) memcpy(dst + (size_t)y * m.cols, m.ptr<uchar>(y), m.cols);
}
// the bordered buffer behind mvImagePyramid[level] (ORBextractor.cc:905-916): (w + 38) x (h + 38)
void ref_orb_level_bordered_copy(void* h, int level, uint8_t* dst) {
    const cv::Mat& m = ((ORB*)h)->mvImagePyramid[level];
    const int W = m.cols + 38, H = m.rows + 38;
    const uchar* base = m.data - 19 * m.step - 19;
    for (int y = 0; y < H; ++y) memcpy(dst + (size_t)y * W, base + (size_t)y * m.step, W);
}
int ref_distribute_octree(const cv::KeyPoint* in, int n, int minX, int maxX, int minY, int maxY, int N, cv::KeyPoint* out, int cap) {
    std::vector<cv::KeyPoint> v(in, in + n);
    std::vector<cv::KeyPoint> r = ORB_SLAM3::ORBextractor::DistributeOctTree(v, minX, maxX, minY, maxY, N, 0);
    for (size_t i = 0; i < r.size() && (int)i < cap; ++i) out[i] = r[i];
    return (int)r.size();
}

// ---- Matcher statics + filters (include/Matcher.hpp:39-92, src/Matcher.cpp) ------------------------------------------------
unsigned ref_hamming(const uint8_t* a, const uint8_t* b) { return HammingDistance(wrap_u8(a, 1, 32, 32), wrap_u8(b, 1, 32, 32)); }
// KnnMatch(vector<Mat>, vector<Mat>, 2): always 2 entries per query
void ref_knn2_firstparty(const uint8_t* q, int nq, const uint8_t* t, int nt, cv::DMatch* out) {
    cv::Mat Q = wrap_u8(q, nq, 32, 32), T = wrap_u8(t, nt, 32, 32);
    MatchResKnn r = Matcher::KnnMatch(rows_of(Q), rows_of(T), 2);
    for (int i = 0; i < nq; ++i) { out[2 * i] = r[i][0]; out[2 * i + 1] = r[i][1]; }
}
// KnnMatch(Mat, Mat, 2) / KnnMatch_cv: min(2, nt) entries per query; returns entries per query. which = 0 KnnMatch, 1 KnnMatch_cv
int ref_knn2_bf(const uint8_t* q, int nq, const uint8_t* t, int nt, cv::DMatch* out, int which) {
    cv::Mat Q = wrap_u8(q, nq, 32, 32), T = wrap_u8(t, nt, 32, 32);
    MatchResKnn r = which ? Matcher::KnnMatch_cv(Q, T, 2) : Matcher::KnnMatch(Q, T, 2);
    int k = nt < 2 ? nt : 2;
    for (int i = 0; i < nq; ++i)
        for (int j = 0; j < 2; ++j) out[2 * i + j] = j < (int)r[i].size() ? r[i][j] : cv::DMatch(i, -1, 0, 0.f);
    return k;
}
int ref_bf_match(const uint8_t* q, int nq, const uint8_t* t, int nt, cv::DMatch* out) {
    MatchRes r = Matcher::BFMatch(wrap_u8(q, nq, 32, 32), wrap_u8(t, nt, 32, 32));
    return put(r, out, nq);
}
// candidate-list 2-NN as every call site builds it: vector<Mat> of row headers in list order -> KnnMatch(vector, vector)
void ref_knn2_candidates(const uint8_t* q, int nq, const uint8_t* t, const int* cand_off, const int* cand_idx, cv::DMatch* out) {
    for (int i = 0; i < nq; ++i) {
        std::vector<cv::Mat> c;
        for (int k = cand_off[i]; k < cand_off[i + 1]; ++k) c.push_back(wrap_u8(t + 32 * (size_t)cand_idx[k], 1, 32, 32));
        MatchResKnn r = Matcher::KnnMatch({wrap_u8(q + 32 * (size_t)i, 1, 32, 32)}, c, 2);
        out[2 * i] = r[0][0]; out[2 * i + 1] = r[0][1];
        out[2 * i].queryIdx = out[2 * i + 1].queryIdx = i;
    }
}
int ref_filter_ratio(const cv::DMatch* knn, int nq, int per, float ratio, cv::DMatch* out) {
    MatchResKnn k;
    for (int i = 0; i < nq; ++i) k.push_back(std::vector<cv::DMatch>(knn + (size_t)i * per, knn + (size_t)(i + 1) * per));
    return put(k.FilterRatio(ratio), out, nq);
}
int ref_filter_threshold(cv::DMatch* m, int n, int th) { MatchRes r; r.assign(m, m + n); r.FilterThreshold(th); return put(r, m, n); }
int ref_filter_orientation(cv::DMatch* m, int n, const cv::KeyPoint* k1, int n1, const cv::KeyPoint* k2, int n2) {
    MatchRes r; r.assign(m, m + n);
    r.FilterOrientation(Keypoints(k1, k1 + n1), Keypoints(k2, k2 + n2));
    return put(r, m, n);
}
int ref_filter_fmatrix(cv::DMatch* m, int n, const cv::KeyPoint* k1, int n1, const cv::KeyPoint* k2, int n2, const float* F12, const float* sigma2, int nlevels) {
    MatchRes r; r.assign(m, m + n);
    cv::Mat F(3, 3, CV_32F);
    for (int i = 0; i < 9; ++i) F.at<float>(i / 3, i % 3) = F12[i];
    r.FilterFMatrix(Keypoints(k1, k1 + n1), Keypoints(k2, k2 + n2), F, std::vector<float>(sigma2, sigma2 + nlevels));
    return put(r, m, n);
}
// Matcher::DBowMatch with the two FeatureVectors given flattened (node ids ascending)
int ref_dbow_match(const uint8_t* d1, int n1, const unsigned* nodes1, const int* off1, const int* idx1, int nn1, const uint8_t* d2, int n2,
                   const unsigned* nodes2, const int* off2, const int* idx2, int nn2, cv::DMatch* out, int cap_pairs) {
    DBoW3::FeatureVector f1, f2;
    for (int i = 0; i < nn1; ++i) for (int k = off1[i]; k < off1[i + 1]; ++k) f1.addFeature(nodes1[i], (unsigned)idx1[k]);
    for (int i = 0; i < nn2; ++i) for (int k = off2[i]; k < off2[i + 1]; ++k) f2.addFeature(nodes2[i], (unsigned)idx2[k]);
    MatchResKnn r = Matcher::DBowMatch(wrap_u8(d1, n1, 32, 32), f1, wrap_u8(d2, n2, 32, 32), f2);
    for (size_t i = 0; i < r.size() && (int)i < cap_pairs; ++i) { out[2 * i] = r[i][0]; out[2 * i + 1] = r[i][1]; }
    return (int)r.size();
}

// ---- Frame: the real constructor (src/Frame.cpp:78-138) ------------------------------------------------------------------
struct RefRig { std::unique_ptr<Pinhole> cam; };
void* ref_rig_create(const char* extractor_yaml, float bf, float baseline) {
    try {
        Frame::extractor_left = ORB(std::string(extractor_yaml));
        Frame::extractor_right = ORB(std::string(extractor_yaml));
        Frame::extractor_wide = ORB(std::string(extractor_yaml));
    } catch (std::exception& e) { fprintf(stderr, "ref_rig_create: %s\n", e.what()); return nullptr; }
    Frame::bf = bf; Frame::b = baseline; Frame::b_2 = baseline / 2;
    Frame::Trl = cv::Mat::eye(4, 4, CV_32F); Frame::Twl = cv::Mat::eye(4, 4, CV_32F);
    RefRig* r = new RefRig;
    const float intr[4] = {955.40503f, 955.40503f, 256.f, 256.f};   // config/camleft.yaml:3
    r->cam.reset(make_cam(intr));
    return r;
}
void ref_rig_destroy(void* r) { delete (RefRig*)r; }
// One three-camera frame through Frame::Frame with an empty optical-flow list (as the reference's tests pass {}). Outputs per
// camera: counts[3], kps / desc [3][cap]; u_right / depth_left [cap]. Returns 0, or -2 when cap is too small.
int ref_rig_frame(void* rr, const uint8_t* imgs, int w, int h, cv::KeyPoint* kps, uint8_t* desc, int* counts, float* u_right, float* depth_left, int cap) {
    RefRig* r = (RefRig*)rr;
    cv::Mat L = wrap_u8(imgs, h, w, w), R = wrap_u8(imgs + (size_t)w * h, h, w, w), W = wrap_u8(imgs + 2 * (size_t)w * h, h, w, w);
    Frame f(L, R, W, 0.0, r->cam.get(), r->cam.get(), r->cam.get(), 0, {});
    ObjectRef objs[3] = {f.LEFT, f.RIGHT, f.WIDE};
    for (int c = 0; c < 3; ++c) {
        int n = (int)objs[c]->kps.size();
        counts[c] = n;
        if (n > cap) return -2;
        if (kps) for (int i = 0; i < n; ++i) kps[(size_t)c * cap + i] = objs[c]->kps[i];
        if (desc) for (int i = 0; i < n; ++i) memcpy(desc + ((size_t)c * cap + i) * 32, objs[c]->desps.ptr<uchar>(i), 32);
    }
    for (int i = 0; i < counts[0]; ++i) { if (u_right) u_right[i] = f.u_right[i]; if (depth_left) depth_left[i] = f.depth_left[i]; }
    return 0;
}
// Frame::ComputeStereoMatch on caller keypoints, against the pyramids of the last ref_rig_frame / ref_rig_extract_lr call
// (extractor_left / extractor_right statics, as the reference reads them).
int ref_rig_extract_lr(void* rr, const uint8_t* left, const uint8_t* right, int w, int h) {
    std::vector<cv::KeyPoint> k; cv::Mat d;
    Frame::extractor_left.Extract(wrap_u8(left, h, w, w), k, d);
    k.clear();
    Frame::extractor_right.Extract(wrap_u8(right, h, w, w), k, d);
    return 0;
}
int ref_rig_stereo(void* rr, const cv::KeyPoint* kl, const uint8_t* dl, int nl, const cv::KeyPoint* kr, const uint8_t* dr, int nr, int w, int h,
                   float* u_right, float* depth_left) {
    RefRig* r = (RefRig*)rr;
    // a Frame without running its constructor's pipeline is not constructible; build one on tiny blank images, then swap in
    // the caller's objects (ComputeStereoMatch only reads left/right kps, desps, bounddingbox and the static extractors' pyramids)
    std::unique_ptr<ObjBox> L(make_obj(&Frame::extractor_left, kl, dl, nl, w, h, nullptr, nullptr, nullptr));
    std::unique_ptr<ObjBox> R(make_obj(&Frame::extractor_right, kr, dr, nr, w, h, nullptr, nullptr, nullptr));
    // Frame has no default constructor; allocate raw storage and construct only the members ComputeStereoMatch touches
    alignas(Frame) static unsigned char raw[sizeof(Frame)];
    Frame* f = reinterpret_cast<Frame*>(raw);
    new (&f->u_right) std::vector<float>();
    new (&f->depth_left) std::vector<float>((size_t)nl, -1.f);
    f->ComputeStereoMatch(L->obj, R->obj);
    for (int i = 0; i < nl; ++i) { u_right[i] = f->u_right[i]; depth_left[i] = f->depth_left[i]; }
    f->u_right.~vector(); f->depth_left.~vector();
    (void)r;
    return 0;
}
// wall seconds for n_frames Frame constructions, sequentially (each one fans its three extractions out over ThreadPool(3) as
// the reference does, src/Frame.cpp:22,118-126). stage_s[3] = accumulated MyTimer "KL", "ORBE", "SMatch" seconds.
double ref_rig_bench(void* rr, const uint8_t* imgs, int n_frames, int w, int h, int repeat, double* stage_s, long long* n_kp) {
    RefRig* r = (RefRig*)rr;
    { std::lock_guard<std::mutex> g(MyTimer::Registry::get().m); MyTimer::Registry::get().acc.clear(); }
    long long kp = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int rep = 0; rep < repeat; ++rep)
        for (int i = 0; i < n_frames; ++i) {
            const uint8_t* p = imgs + (size_t)i * 3 * w * h;
            Frame f(wrap_u8(p, h, w, w), wrap_u8(p + (size_t)w * h, h, w, w), wrap_u8(p + 2 * (size_t)w * h, h, w, w), 0.0, r->cam.get(), r->cam.get(),
                    r->cam.get(), (uint)i, {});
            kp += (long long)f.LEFT->size() + f.RIGHT->size() + f.WIDE->size();
        }
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (stage_s) { auto& a = MyTimer::Registry::get().acc; stage_s[0] = a["KL"].first; stage_s[1] = a["ORBE"].first; stage_s[2] = a["SMatch"].first; }
    if (n_kp) *n_kp = kp;
    return s;
}

// ---- Object: grid, projection, BoW ---------------------------------------------------------------------------------------
void* ref_obj_create(void* orb, const cv::KeyPoint* kps, const uint8_t* desc, int n, int w, int h, const float* intr, const float* Rcw, const float* tcw) {
    return make_obj((ORB*)orb, kps, desc, n, w, h, intr, Rcw, tcw);
}
void ref_obj_destroy(void* o) { delete (ObjBox*)o; }
int ref_obj_features_in_area(void* o, float x, float y, float r, int* out, int cap) {
    std::vector<size_t> v = ((ObjBox*)o)->obj->GetFeaturesInArea(x, y, r);
    for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
    return (int)v.size();
}
// Object::ProjectBunchMapPoints, one MapPoint per call in the caller's order so that each point's own outcome is observable
// (the set overload iterates an unordered_set — order is the caller's by contract). out_idx[m] = index handed to AddMapPoint or -1.
int ref_project_match(void* o, const float* mp_xyz, const uint8_t* mp_desc, const int* mp_level, int n_mp, float r_threshold, int* out_idx) {
    ObjectRef obj = ((ObjBox*)o)->obj;
    int cnt = 0;
    for (int m = 0; m < n_mp; ++m) {
        MapPointRef mp = make_mp(mp_xyz + 3 * m, mp_desc + 32 * (size_t)m, mp_level[m], (uint)m);
        uint c = obj->ProjectBunchMapPoints(std::vector<MapPointRef>{mp}, r_threshold);
        out_idx[m] = c ? (int)obj->GetMapPointIdx(mp) : -1;
        cnt += (int)c;
        obj->clear();
    }
    return cnt;
}
// DBoW3 vocabulary through DBoW3's own binary loader (Vocabulary::load -> fromStream), then Object::ComputeBow.
int ref_voc_load(const char* path) { try { Object::voc.load(std::string(path)); } catch (std::exception& e) { fprintf(stderr, "ref_voc_load: %s\n", e.what()); return -1; } catch (std::string& e) { fprintf(stderr, "ref_voc_load: %s\n", e.c_str()); return -1; } return (int)Object::voc.size(); }
// the loaded vocabulary written back by DBoW3's own Vocabulary::save as the UNCOMPRESSED binary stream (toStream): how the test
// side gets at the shipped, quicklz-compressed Vocabulary/orbvoc.dbow3 without a second decompressor
int ref_voc_save(const char* path) { try { Object::voc.save(std::string(path), false); } catch (std::exception& e) { fprintf(stderr, "ref_voc_save: %s\n", e.what()); return -1; } catch (std::string& e) { fprintf(stderr, "ref_voc_save: %s\n", e.c_str()); return -1; } return 0; }
// returns BowVector size; bow_ids / bow_vals [<= n]; FeatureVector flattened: fv_nodes, fv_off [n_fv + 1], fv_idx
int ref_obj_compute_bow(void* o, unsigned* bow_ids, double* bow_vals, unsigned* fv_nodes, int* fv_off, int* fv_idx, int* n_fv) {
    ObjectRef obj = ((ObjBox*)o)->obj;
    obj->is_bowed = false; obj->bow_vector.clear(); obj->bow_feature.clear();
    obj->ComputeBow();
    int i = 0;
    for (auto& p : obj->bow_vector) { bow_ids[i] = p.first; bow_vals[i] = p.second; ++i; }
    int k = 0, nf = 0;
    fv_off[0] = 0;
    for (auto& p : obj->bow_feature) { fv_nodes[nf] = p.first; for (auto f : p.second) fv_idx[k++] = (int)f; fv_off[++nf] = k; }
    *n_fv = nf;
    return i;
}

// ---- MapPoint::ComputeDistinctiveDescriptors -----------------------------------------------------------------------------
// desc rows [off[m], off[m+1]) = the observed descriptors of point m; every observation lives in its own Object / KeyFrame key.
// The order in which the reference walks its unordered containers is read back (order_out: observation row per position) so
// that the caller can present the same order to the checker; best_pos = BestIdx within that order, out_desc = MapPoint::desp after.
void ref_distinctive(void* orb, const uint8_t* desc, const int* off, int n_mp, int* best_row, uint8_t* out_desc) {
    for (int m = 0; m < n_mp; ++m) {
        const int N = off[m + 1] - off[m];
        uint8_t zero[32] = {0};
        MapPointRef mp = make_mp(kUnitZ, zero, 0, (uint)m);
        std::vector<std::unique_ptr<ObjBox>> objs;
        for (int i = 0; i < N; ++i) {
            cv::KeyPoint kp(10.f, 10.f, 31.f);
            objs.emplace_back(make_obj((ORB*)orb, &kp, desc + 32 * (size_t)(off[m] + i), 1, 64, 64, nullptr, nullptr, nullptr));
            objs.back()->obj->AddMapPoint(mp, 0);
            mp->BindKeyFrame((KeyFrame)(uintptr_t)(0x1000 + 16 * i), objs.back()->obj);
        }
        mp->ComputeDistinctiveDescriptors();
        cv::Mat d = mp->GetDesp();
        best_row[m] = -1;
        if (N > 0) {
            memcpy(out_desc + 32 * (size_t)m, d.ptr<uchar>(0), 32);
            // which observation was cloned: walk the observations in the reference's own iteration order, first byte-equal row
            for (auto& p : mp->GetAllObservation()) {
                for (auto& ob : p.second) {
                    if (best_row[m] < 0 && memcmp(ob->desps.ptr<uchar>(0), d.ptr<uchar>(0), 32) == 0) {
                        for (int i = 0; i < N; ++i) if (objs[i]->obj == ob) best_row[m] = i;
                    }
                }
            }
        } else memset(out_desc + 32 * (size_t)m, 0, 32);
    }
}
// The iteration order the reference uses for point m's observations (row indices), for presenting the same order to the oracle.
void ref_distinctive_order(int N, int* order) {
    std::unordered_map<KeyFrame, int> keys;
    for (int i = 0; i < N; ++i) keys[(KeyFrame)(uintptr_t)(0x1000 + 16 * i)] = i;
    int k = 0;
    for (auto& p : keys) order[k++] = p.second;
}

// ---- KL_Track (src/Frame.cpp:34-76) -------------------------------------------------------------------------------------
// obj1 = prev image with n keypoints each owning a MapPoint; obj2 = next image without keypoints. Outputs in the reference's
// GetMapPointsVector() order: order[i] = input keypoint of position i; new_kps = obj2->kps afterwards (cnt entries) and
// src[j] = input keypoint index the j-th new keypoint came from. Returns cnt.
int ref_kl_track(void* orb, const uint8_t* prev, const uint8_t* next, int w, int h, int stride, const cv::KeyPoint* kps, int n, cv::KeyPoint* new_kps, int* src) {
    const float intr[4] = {500.f, 500.f, w / 2.f, h / 2.f};
    std::unique_ptr<Pinhole> cam(make_cam(intr));
    ObjectRef o1 = std::make_shared<Object>(cam.get(), wrap_u8(prev, h, w, stride), (ORB*)orb, CAM_NAME::L);
    ObjectRef o2 = std::make_shared<Object>(cam.get(), wrap_u8(next, h, w, stride), (ORB*)orb, CAM_NAME::L);
    o1->kps.assign(kps, kps + n);
    uint8_t zero[32] = {0};
    std::vector<MapPointRef> mps;
    for (int i = 0; i < n; ++i) { mps.push_back(make_mp(kUnitZ, zero, 0, (uint)i)); o1->AddMapPoint(mps.back(), i); }
    std::unordered_map<MapPointRef, uint> mp2idx;
    uint cnt = KL_Track(o1, o2, mp2idx);
    for (auto& p : mp2idx) { new_kps[p.second] = o2->kps[p.second]; src[p.second] = (int)p.first->id; }
    return (int)cnt;
}

// ---- Map::Fuse / ComputeF12, Tracker::Wnd_Track ---------------------------------------------------------------------------
// Map::Fuse, one MapPoint per call (see ref_project_match). depth_left = KeyFrame::depth_left, bf = KeyFrame::bf (static).
int ref_fuse_match(void* o, const char* system_yaml, const float* depth_left, int n, float bf, const float* mp_xyz, const float* mp_normal, const uint8_t* mp_desc,
                   const int* mp_level, int n_mp, int* out_idx) {
    ObjectRef obj = ((ObjBox*)o)->obj;
    Map map{std::string(system_yaml)};
    alignas(Frame) static unsigned char raw[sizeof(Frame)];
    Frame* kf = reinterpret_cast<Frame*>(raw);
    new (&kf->depth_left) std::vector<float>(depth_left, depth_left + n);
    Frame::bf = bf;
    int cnt = 0;
    for (int m = 0; m < n_mp; ++m) {
        MapPointRef mp = make_mp(mp_xyz + 3 * m, mp_desc + 32 * (size_t)m, mp_level[m], (uint)m);
        mp->norm_vec = (cv::Mat_<float>(3, 1) << mp_normal[3 * m], mp_normal[3 * m + 1], mp_normal[3 * m + 2]);
        int c = map.Fuse(obj, kf, std::unordered_set<MapPointRef>{mp});
        out_idx[m] = c ? (int)obj->GetMapPointIdx(mp) : -1;
        cnt += c;
        obj->clear();
    }
    kf->depth_left.~vector();
    return cnt;
}
void ref_compute_f12(void* o1, void* o2, float* F12) {
    cv::Mat F = Map::ComputeF12(((ObjBox*)o1)->obj, ((ObjBox*)o2)->obj);
    for (int i = 0; i < 9; ++i) F12[i] = F.at<float>(i / 3, i % 3);
}
// Tracker::Wnd_Track, one owning keypoint per call. out_idx[q] = index handed to obj2->AddMapPoint, or -1.
int ref_wnd_track(void* o1, void* o2, const char* system_yaml, const int* q_idx, int n_q, int* out_idx) {
    ObjectRef a = ((ObjBox*)o1)->obj, b = ((ObjBox*)o2)->obj;
    Map map{std::string(system_yaml)};
    Tracker trk(&map, nullptr, std::string(system_yaml));
    uint8_t zero[32] = {0};
    int cnt = 0;
    for (int q = 0; q < n_q; ++q) {
        MapPointRef mp = make_mp(kUnitZ, zero, 0, (uint)q);
        a->AddMapPoint(mp, (size_t)q_idx[q]);
        uint c = trk.Wnd_Track(a, b);
        out_idx[q] = c ? (int)b->GetMapPointIdx(mp) : -1;
        cnt += (int)c;
        a->clear(); b->clear();
    }
    return cnt;
}

}  // extern "C"
