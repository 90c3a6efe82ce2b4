// C++ host mirror of the reference's extractor / Matcher / Frame surface on top of the C ABI (include/mcv_b200.h).
//
// Same class names, method names, argument meaning and result layout as the reference, so code written against
//   modules/local_feature/BaseExtractor/BaseExtractor.hpp:10-18      (Keypoints, Desps, BaseExtractor::Extract)
//   modules/local_feature/ORB/ORBExtractor.hpp:8-18                   (MCVSLAM::ORB)
//   modules/local_feature/ORB/orb3_extractor/ORBextractor.h:44-99     (ORB_SLAM3::ORBextractor public surface)
//   include/Matcher.hpp:14-92                                         (HammingDistance, MatchRes, MatchResKnn, Matcher)
//   include/Frame.hpp:25,36-51 / src/Frame.cpp:78-138,150-328         (Frame ORBE + SMatch stages, u_right, depth_left)
//   include/Object.hpp:167-197 / src/Object.cpp:208-236               (Object::ProjectBunchMapPoints)
// keeps compiling, while every byte of compute happens in libmcv_b200.so (sm_100a kernels). Header-only; nothing here
// computes on the CPU except the reference's own order-defining O(matches) filter epilogues, which are host code inside
// the library too. There is no CPU fallback: without a CUDA device the constructors / calls report MCV_ERR_NO_DEVICE.
//
// Differences from the reference, all deliberate:
//   * ORB::Extract never spins forever on failure (ORBExtractor.cpp:32-36); it returns the negative mcv_status and
//     ORB::last_error() holds the text. An empty image still returns -1 (ORBextractor.cc:834).
//   * mvImagePyramid is filled from the device pyramid after each Extract only while keep_host_pyramid is true (default);
//     the stereo and rig paths read the device copy and do not need it.
//   * Object::ProjectBunchMapPoints takes an ORDERED vector of MapPoint views: the reference iterates an unordered_set of
//     shared_ptrs (pointer-hash order), which no drop-in can reproduce; per-MapPoint results are order-independent, only
//     the last-writer-wins AddMapPoint bookkeeping is, and that stays with the caller.
#pragma once
#include <stdio.h>

#include <fstream>
#include <future>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mcv_b200.h"
#include "cv_shim.hpp"

static_assert(sizeof(cv::KeyPoint) == sizeof(mcv_keypoint) && sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
static_assert(sizeof(cv::DMatch) == sizeof(mcv_dmatch) && sizeof(cv::DMatch) == 16, "cv::DMatch layout");

#define ORB_GOOD_THRESHOLD 46  // include/Matcher.hpp:14

namespace mcv_host {

inline mcv_keypoint* kp_ptr(std::vector<cv::KeyPoint>& v) { return reinterpret_cast<mcv_keypoint*>(v.data()); }
inline const mcv_keypoint* kp_ptr(const std::vector<cv::KeyPoint>& v) { return reinterpret_cast<const mcv_keypoint*>(v.data()); }
inline mcv_dmatch* dm_ptr(std::vector<cv::DMatch>& v) { return reinterpret_cast<mcv_dmatch*>(v.data()); }

// The flat `key: value` subset of pyp::yaml that config/extractor.yaml and config/frame.yaml use, incl. ${CURRENT_FOLDER}.
inline std::map<std::string, std::string> parse_flat_yaml(const std::string& path) {
    std::map<std::string, std::string> kv;
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open config " + path);
    const size_t slash = path.find_last_of('/');
    const std::string folder = slash == std::string::npos ? "." : path.substr(0, slash);
    auto trim = [](std::string s) {
        const char* ws = " \t\r\n\"";
        const size_t b = s.find_first_not_of(ws), e = s.find_last_not_of(ws);
        return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
    };
    std::string line;
    while (std::getline(f, line)) {
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        const size_t colon = line.find(':');
        if (colon == std::string::npos) continue;
        std::string v = trim(line.substr(colon + 1));
        const size_t cf = v.find("${CURRENT_FOLDER}");
        if (cf != std::string::npos) v.replace(cf, 17, folder);
        kv[trim(line.substr(0, colon))] = v;
    }
    return kv;
}

}  // namespace mcv_host

// =========================================================================================================
namespace ORB_SLAM3 {

class ORBextractor {
   public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
    ORBextractor() {}
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) { init(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST); }
    ~ORBextractor() {}

    // ORBextractor::init (ORBextractor.cc:407-457). device: CUDA ordinal (the reference has no such notion; default 0).
    void init(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST, int device = 0) {
        nfeatures = _nfeatures; scaleFactor = _scaleFactor; nlevels = _nlevels; iniThFAST = _iniThFAST; minThFAST = _minThFAST;
        mcv_orb_params p{nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST};
        mcv_orb* raw = nullptr;
        status_ = mcv_orb_create(&p, device, nullptr, &raw);
        h_.reset(raw, [](mcv_orb* h) { mcv_orb_destroy(h); });
        if (status_ != MCV_OK) { err_ = mcv_last_error(); return; }
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        mnFeaturesPerLevel.resize(nlevels);
        mcv_orb_get_scales(raw, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(), mnFeaturesPerLevel.data());
    }

    // ORBextractor::operator() (ORBextractor.cc:831-899). _keypoints is in/out: pre-seeded keypoints are appended per
    // octave after the quadtree keypoints. Mask and vLappingArea are ignored, as in the reference.
    int operator()(const cv::Mat& _image, const cv::Mat& /*_mask*/, std::vector<cv::KeyPoint>& _keypoints, cv::Mat& _descriptors,
                   std::vector<int>& /*vLappingArea*/) {
        if (_image.empty()) return -1;
        if (!h_) { status_ = MCV_ERR_BAD_ARG; err_ = "extractor not initialised"; return status_; }
        const std::vector<cv::KeyPoint> seeds = _keypoints;
        const int cap = mcv_orb_max_keypoints_for(h_.get(), _image.cols, _image.rows, (int)seeds.size());
        std::vector<cv::KeyPoint> out((size_t)cap);
        cv::Mat desc(cap, 32, CV_8U);
        int n = 0;
        status_ = mcv_orb_extract(h_.get(), _image.data, _image.cols, _image.rows, _image.step, mcv_host::kp_ptr(seeds), (int)seeds.size(),
                                  mcv_host::kp_ptr(out), desc.data, cap, &n);
        if (status_ != MCV_OK) { err_ = mcv_last_error(); return status_ == MCV_ERR_EMPTY_IMAGE ? -1 : status_; }
        out.resize((size_t)n);
        _keypoints.swap(out);
        _descriptors.create(n, 32, CV_8U);  // continuous N x 32 CV_8U (ORBextractor.cc:857)
        if (n) memcpy(_descriptors.data, desc.data, (size_t)n * 32);
        if (keep_host_pyramid) {
            mvImagePyramid.resize(nlevels);
            for (int l = 0; l < nlevels; ++l) {
                int w = 0, h = 0;
                mcv_orb_download_level(h_.get(), 0, l, nullptr, 0, &w, &h);
                mvImagePyramid[l].create(h, w, CV_8U);
                mcv_orb_download_level(h_.get(), 0, l, mvImagePyramid[l].data, mvImagePyramid[l].step, &w, &h);
            }
        }
        return n;
    }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    std::vector<cv::Mat> mvImagePyramid;
    bool keep_host_pyramid = true;

    // static DistributeOctTree (ORBextractor.h:73-74, ORBextractor.cc:524-580); `level` is unused, as in the reference.
    static std::vector<cv::KeyPoint> DistributeOctTree(const std::vector<cv::KeyPoint>& vToDistributeKeys, const int& minX, const int& maxX,
                                                       const int& minY, const int& maxY, const int& nFeatures, const int& /*level*/) {
        static thread_local std::shared_ptr<mcv_orb> scratch;
        if (!scratch) {
            mcv_orb_params p{1000, 1.2f, 1, 20, 7};
            mcv_orb* raw = nullptr;
            if (mcv_orb_create(&p, 0, nullptr, &raw) != MCV_OK) throw std::runtime_error(std::string("DistributeOctTree: ") + mcv_last_error());
            scratch.reset(raw, [](mcv_orb* h) { mcv_orb_destroy(h); });
        }
        std::vector<cv::KeyPoint> out(vToDistributeKeys.size() + 16);
        int n = 0;
        const mcv_status st = mcv_orb_distribute_octree(scratch.get(), mcv_host::kp_ptr(vToDistributeKeys), (int)vToDistributeKeys.size(), minX, maxX,
                                                        minY, maxY, nFeatures, mcv_host::kp_ptr(out), (int)out.size(), &n);
        if (st != MCV_OK) throw std::runtime_error(std::string("DistributeOctTree: ") + mcv_last_error());
        out.resize((size_t)n);
        return out;
    }

    mcv_orb* handle() const { return h_.get(); }
    mcv_status status() const { return status_; }
    const std::string& last_error() const { return err_; }

   protected:
    int nfeatures = 0;
    float scaleFactor = 1.2f;
    int nlevels = 0;
    int iniThFAST = 0;
    int minThFAST = 0;
    std::vector<int> mnFeaturesPerLevel;
    std::shared_ptr<mcv_orb> h_;  // copies share one engine handle (`extractor_left = ORB(path)`, src/Frame.cpp:359-361)
    mcv_status status_ = MCV_OK;
    std::string err_;

   public:
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;
};

}  // namespace ORB_SLAM3

// =========================================================================================================
namespace MCVSLAM {

using Keypoints = std::vector<cv::KeyPoint>;
using Desps = cv::Mat;
typedef unsigned int uint;

class BaseExtractor {
   public:
    BaseExtractor() {}
    virtual ~BaseExtractor() {}
    virtual int Extract(const cv::Mat img, Keypoints& kps, Desps& desps) = 0;
};

class ORB : public MCVSLAM::BaseExtractor, public ORB_SLAM3::ORBextractor {
   public:
    ORB() {}
    ORB(const std::string& config_path) {
        Parse(config_path);
        init(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
    }
    ~ORB() {}

    // ORBExtractor.cpp:27-38. Returns the keypoint count, -1 for an empty image, or a negative mcv_status (<= -2).
    int Extract(const cv::Mat img, Keypoints& kps, Desps& desps) override {
        std::vector<int> lap = {0, 0};
        return (*this)(img, cv::Mat(), kps, desps, lap);
    }

   private:
    void Parse(const std::string& config_file) {  // ORBExtractor.cpp:10-18
        auto root = mcv_host::parse_flat_yaml(config_file);
        nfeatures = std::stoi(root.at("nkeypoints"));
        iniThFAST = std::stoi(root.at("ORBextractor.iniThFAST"));
        minThFAST = std::stoi(root.at("ORBextractor.minThFAST"));
        scaleFactor = std::stof(root.at("scale_factor"));
        nlevels = std::stoi(root.at("nlevels"));
    }
};

enum MATCH_DISTANCE { HAMMING = 0, COS = 1, NORM2 = 2 };

// include/Matcher.hpp:19-33 — the one scalar helper of the path that callers use directly on two rows; it is the
// definition the kernels' __popc sums are tested against, not a compute path of its own.
static inline uint HammingDistance(const cv::Mat& a, const cv::Mat& b) {
    const uint32_t* pa = a.ptr<uint32_t>();
    const uint32_t* pb = b.ptr<uint32_t>();
    uint dist = 0;
    for (int i = 0; i < 8; i++) dist += (uint)__builtin_popcount(pa[i] ^ pb[i]);
    return dist;
}

class MatchRes : public std::vector<cv::DMatch> {
   public:
    MatchRes& FilterThreshold(const int thres_hold = ORB_GOOD_THRESHOLD) {  // src/Matcher.cpp:23-35
        int n = (int)size();
        mcv_filter_threshold(mcv_host::dm_ptr(*this), &n, thres_hold);
        resize((size_t)n);
        return *this;
    }
    MatchRes& FilterOrientation(const Keypoints& kps1, const Keypoints& kps2) {  // src/Matcher.cpp:44-74
        int n = (int)size();
        mcv_filter_orientation(mcv_host::dm_ptr(*this), &n, mcv_host::kp_ptr(kps1), (int)kps1.size(), mcv_host::kp_ptr(kps2), (int)kps2.size());
        resize((size_t)n);
        return *this;
    }
    // src/Matcher.cpp:76-91; F12 = 9 floats row-major (the reference passes a 3x3 CV_32F cv::Mat)
    MatchRes& FilterFMatrix(const Keypoints& kps1, const Keypoints& kps2, const float* F12, const std::vector<float>& LevelSigma2) {
        int n = (int)size();
        mcv_filter_fmatrix(mcv_host::dm_ptr(*this), &n, mcv_host::kp_ptr(kps1), (int)kps1.size(), mcv_host::kp_ptr(kps2), (int)kps2.size(), F12,
                           LevelSigma2.data(), (int)LevelSigma2.size());
        resize((size_t)n);
        return *this;
    }
};

class MatchResKnn : public std::vector<std::vector<cv::DMatch>> {
   public:
    MatchRes FilterRatio(const float ratio = 0.6) {  // src/Matcher.cpp:100-111
        MatchRes ret;
        for (auto& v : *this) {
            if (v.empty()) continue;
            int n = 0;
            cv::DMatch keep;
            mcv_filter_ratio(reinterpret_cast<const mcv_dmatch*>(v.data()), 1, (int)v.size() >= 2 ? 2 : 1, ratio, reinterpret_cast<mcv_dmatch*>(&keep), &n);
            if (n) ret.push_back(keep);
        }
        return ret;
    }
};

class Matcher {
   public:
    static Matcher& GetInstance(MATCH_DISTANCE _dis_mode) {
        static Matcher m;
        m.dis_mode = _dis_mode;
        return m;
    }

    // KnnMatch(const cv::Mat&, const cv::Mat&, k) == cv::BFMatcher(NORM_HAMMING).knnMatch (src/Matcher.cpp:134-138,304-308)
    static MatchResKnn KnnMatch(const cv::Mat& desp1, const cv::Mat& desp2, int k = 2, uint /*n_threads*/ = 0) {
        check_k(k);
        const cv::Mat q = contiguous(desp1), t = contiguous(desp2);
        std::vector<cv::DMatch> flat((size_t)q.rows * 2);
        int kk = 0;
        check(mcv_knn2_bf(q.data, q.rows, t.data, t.rows, mcv_host::dm_ptr(flat), &kk));
        return unflatten(flat, q.rows, kk);
    }
    static MatchResKnn KnnMatch_cv(const cv::Mat& desp1, const cv::Mat& desp2, int k = 2) { return KnnMatch(desp1, desp2, k); }

    // KnnMatch(vector<Mat>, vector<Mat>, k) (src/Matcher.cpp:245-302): always two entries per query, (0, 999) padding.
    static MatchResKnn KnnMatch(const std::vector<cv::Mat>& desp1, const std::vector<cv::Mat>& desp2, int k = 2, uint /*n_threads*/ = 0) {
        check_k(k);
        const std::vector<uint8_t> q = gather_rows(desp1), t = gather_rows(desp2);
        std::vector<cv::DMatch> flat(desp1.size() * 2);
        check(mcv_knn2_firstparty(q.data(), (int)desp1.size(), t.data(), (int)desp2.size(), mcv_host::dm_ptr(flat)));
        return unflatten(flat, (int)desp1.size(), 2);
    }
    // the keypoint-checked overload is a stub in the reference (returns an empty result, src/Matcher.cpp:195-242)
    static MatchResKnn KnnMatch(const std::vector<cv::KeyPoint>&, const std::vector<cv::Mat>&, const std::vector<cv::KeyPoint>&, const std::vector<cv::Mat>&,
                                const std::vector<float>, int = 2, uint = 0) { return MatchResKnn(); }

    static MatchRes BFMatch(const Desps& desp1, const Desps& desp2) {  // src/Matcher.cpp:140-144
        const cv::Mat q = contiguous(desp1), t = contiguous(desp2);
        MatchRes r;
        r.resize((size_t)q.rows);
        check(mcv_bf_match(q.data, q.rows, t.data, t.rows, mcv_host::dm_ptr(r)));
        return r;
    }

    // DBowMatch (src/Matcher.cpp:146-193). FeatureVector = std::map<node id, std::vector<feature index>> (DBoW3::FeatureVector).
    typedef std::map<unsigned int, std::vector<unsigned int>> FeatureVector;
    static MatchResKnn DBowMatch(const Desps& desp1, const FeatureVector& bow_feat1, const Desps& desp2, const FeatureVector& bow_feat2) {
        const cv::Mat d1 = contiguous(desp1), d2 = contiguous(desp2);
        std::vector<uint32_t> ids1, ids2;
        std::vector<int32_t> off1, off2, idx1, idx2;
        flatten(bow_feat1, ids1, off1, idx1);
        flatten(bow_feat2, ids2, off2, idx2);
        std::vector<cv::DMatch> flat((size_t)std::max(d1.rows, 1) * 2);
        int pairs = 0;
        check(mcv_dbow_match(d1.data, d1.rows, ids1.data(), off1.data(), idx1.data(), (int)ids1.size(), d2.data, d2.rows, ids2.data(), off2.data(),
                             idx2.data(), (int)ids2.size(), mcv_host::dm_ptr(flat), &pairs));
        return unflatten(flat, pairs, 2);
    }

    MATCH_DISTANCE dis_mode = HAMMING;

   private:
    Matcher() {}
    ~Matcher() {}
    static void check(mcv_status st) { if (st != MCV_OK) throw std::runtime_error(std::string("Matcher: ") + mcv_last_error() + " (status " + std::to_string(st) + ")"); }
    static void check_k(int k) { if (k != 2) throw std::invalid_argument("Matcher: only k = 2 is used by the reference and supported"); }
    static cv::Mat contiguous(const cv::Mat& m) { return (m.empty() || m.isContinuous()) ? m : m.clone(); }
    static std::vector<uint8_t> gather_rows(const std::vector<cv::Mat>& rows) {
        std::vector<uint8_t> buf(rows.size() * 32 + 32);
        for (size_t i = 0; i < rows.size(); ++i) memcpy(buf.data() + i * 32, rows[i].data, 32);
        return buf;
    }
    static MatchResKnn unflatten(const std::vector<cv::DMatch>& flat, int nq, int per) {
        MatchResKnn r;
        r.resize((size_t)nq);
        for (int i = 0; i < nq; ++i) r[i].assign(flat.begin() + (size_t)2 * i, flat.begin() + (size_t)2 * i + per);
        return r;
    }
    static void flatten(const FeatureVector& fv, std::vector<uint32_t>& ids, std::vector<int32_t>& off, std::vector<int32_t>& idx) {
        off.push_back(0);
        for (const auto& kv : fv) {
            ids.push_back(kv.first);
            for (unsigned int f : kv.second) idx.push_back((int32_t)f);
            off.push_back((int32_t)idx.size());
        }
        if (idx.empty()) idx.push_back(0);
        if (ids.empty()) ids.push_back(0);
    }
};

// ---------------------------------------------------------------------------------------------------------
// Object / Frame: the callers either side of the path, reduced to what the path reads and writes.
// ---------------------------------------------------------------------------------------------------------
}  // namespace MCVSLAM

// DBoW3 as far as Object::ComputeBow / Matcher::DBowMatch use it (modules/DBow3/src/{BowVector,FeatureVector,Vocabulary}.h)
namespace DBoW3 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
enum WeightingType { TF_IDF, TF, IDF, BINARY };
enum LNorm { L1, L2 };
class BowVector : public std::map<WordId, WordValue> {};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {};

// The vocabulary tree lives in device memory; it is built from the flat arrays a maintainer gets by walking the real
// DBoW3::Vocabulary's nodes once after load() (INTEGRATION.md shows the loop).
class Vocabulary {
   public:
    Vocabulary() {}
    Vocabulary(int n_nodes, const int32_t* child_off, const uint32_t* child_ids, const uint8_t* node_desc, const int32_t* word_id, const double* weight,
               int L, WeightingType weighting, int norm /* 0 none, 1 L1, 2 L2 */, int device = 0) {
        mcv_voc* raw = nullptr;
        if (mcv_voc_create(n_nodes, child_off, child_ids, node_desc, word_id, weight, L, (int)weighting, norm, device, &raw) != MCV_OK)
            throw std::runtime_error(std::string("Vocabulary: ") + mcv_last_error());
        h.reset(raw, mcv_voc_destroy);
    }
    bool empty() const { return !h; }
    // void transform(const std::vector<cv::Mat>& features, BowVector& v, FeatureVector& fv, int levelsup) const — Vocabulary.cpp:572-633
    void transform(const std::vector<cv::Mat>& features, BowVector& v, FeatureVector& fv, int levelsup) const {
        v.clear(); fv.clear();
        if (empty()) return;
        const int n = (int)features.size();
        std::vector<uint8_t> d((size_t)n * 32 + 32);
        for (int i = 0; i < n; ++i) memcpy(&d[(size_t)i * 32], features[i].data, 32);
        std::vector<uint32_t> bi((size_t)n + 1), fn((size_t)n + 1);
        std::vector<double> bv((size_t)n + 1);
        std::vector<int32_t> fo((size_t)n + 2), fi((size_t)n + 1);
        int nb = 0, nf = 0;
        if (mcv_bow_transform(h.get(), d.data(), n, levelsup, nullptr, nullptr, nullptr, bi.data(), bv.data(), &nb, fn.data(), fo.data(), fi.data(), &nf) != MCV_OK)
            throw std::runtime_error(std::string("Vocabulary::transform: ") + mcv_last_error());
        for (int k = 0; k < nb; ++k) v.insert(v.end(), std::make_pair(bi[k], bv[k]));
        for (int m = 0; m < nf; ++m) fv.insert(fv.end(), std::make_pair(fn[m], std::vector<unsigned int>(fi.begin() + fo[m], fi.begin() + fo[m + 1])));
    }

   private:
    std::shared_ptr<mcv_voc> h;
};
}  // namespace DBoW3

namespace MCVSLAM {

struct MapPointView {      // what ProjectBunchMapPoints / Map::Fuse read of a MapPoint: GetWorldPos(), GetDesp(), level, GetNormalVector()
    float xyz[3];
    uint8_t desp[32];
    int level;
    float normal[3] = {0, 0, 0};
};

class Object {
   public:
    Object() {}
    Object(const cv::Mat& _img, ORB* _extractor) : img(_img), extractor(_extractor) {}
    size_t size() const { return kps.size(); }

    // Rcw 3x3 row-major, tcw 3, pinhole intrinsics (modules/camera/Pinhole.cpp:45-47)
    void SetPose(const float* R, const float* t) { memcpy(Rcw, R, sizeof(Rcw)); memcpy(tcw, t, sizeof(tcw)); }
    void SetIntrinsics(float fx, float fy, float cx, float cy) { intr[0] = fx; intr[1] = fy; intr[2] = cx; intr[3] = cy; }

    // Object::ProjectBunchMapPoints (src/Object.cpp:208-236) incl. AssignFeaturesToGrid / GetFeaturesInArea
    // (src/Object.cpp:182-201,249-308). matched_idx[m] = keypoint index the reference would pass to AddMapPoint, or -1.
    uint ProjectBunchMapPoints(const std::vector<MapPointView>& mps, float r_threshold, std::vector<int>& matched_idx, std::vector<int>* matched_dist = nullptr) {
        const int n_mp = (int)mps.size();
        std::vector<float> xyz((size_t)n_mp * 3 + 3);
        std::vector<uint8_t> md((size_t)n_mp * 32 + 32);
        std::vector<int32_t> lvl((size_t)n_mp + 1), od((size_t)n_mp + 1);
        for (int m = 0; m < n_mp; ++m) { memcpy(&xyz[3 * (size_t)m], mps[m].xyz, 12); memcpy(&md[32 * (size_t)m], mps[m].desp, 32); lvl[m] = mps[m].level; }
        matched_idx.assign((size_t)n_mp, -1);
        int cnt = 0;
        const cv::Mat d = (desps.empty() || desps.isContinuous()) ? desps : desps.clone();
        const mcv_status st = mcv_project_match(mcv_host::kp_ptr(kps), d.data, (int)kps.size(), img.cols, img.rows, extractor->mvScaleFactor.data(),
                                                (int)extractor->mvScaleFactor.size(), Rcw, tcw, intr, xyz.data(), md.data(), lvl.data(), n_mp, r_threshold,
                                                matched_idx.data(), od.data(), &cnt);
        if (st != MCV_OK) throw std::runtime_error(std::string("ProjectBunchMapPoints: ") + mcv_last_error());
        if (matched_dist) matched_dist->assign(od.begin(), od.begin() + n_mp);
        return (uint)cnt;
    }

    // Matching front-end of Map::Fuse(obj1 = this, kf, mps) (src/Map.cpp:478-527) for an ORDERED vector of MapPoint views:
    // matched_idx[m] = the keypoint index the reference hands to AddMapPoint / ReplaceMappoint (best_idx), or -1. The map
    // surgery (src/Map.cpp:528-547) stays with the caller. kf_depth_left / kf_bf = kf->depth_left, kf->bf.
    uint FuseMatch(const std::vector<MapPointView>& mps, const std::vector<float>& kf_depth_left, float kf_bf, std::vector<int>& matched_idx,
                   std::vector<int>* matched_dist = nullptr) {
        const int n_mp = (int)mps.size();
        std::vector<float> xyz((size_t)n_mp * 3 + 3), nrm((size_t)n_mp * 3 + 3);
        std::vector<uint8_t> md((size_t)n_mp * 32 + 32);
        std::vector<int32_t> lvl((size_t)n_mp + 1), od((size_t)n_mp + 1);
        for (int m = 0; m < n_mp; ++m) {
            memcpy(&xyz[3 * (size_t)m], mps[m].xyz, 12); memcpy(&nrm[3 * (size_t)m], mps[m].normal, 12);
            memcpy(&md[32 * (size_t)m], mps[m].desp, 32); lvl[m] = mps[m].level;
        }
        if (kf_depth_left.size() != kps.size()) throw std::runtime_error("FuseMatch: depth_left must have one entry per keypoint");
        matched_idx.assign((size_t)n_mp, -1);
        float Ow[3];   // mOw = -mRwc * mtcw (src/Object.cpp:174)
        for (int r = 0; r < 3; ++r) Ow[r] = -((Rcw[r] * tcw[0] + Rcw[3 + r] * tcw[1]) + Rcw[6 + r] * tcw[2]);
        int cnt = 0;
        const cv::Mat d = (desps.empty() || desps.isContinuous()) ? desps : desps.clone();
        const mcv_status st = mcv_fuse_match(mcv_host::kp_ptr(kps), d.data, (int)kps.size(), img.cols, img.rows, extractor->mvLevelSigma2.data(),
                                             extractor->mvInvLevelSigma2.data(), (int)extractor->mvLevelSigma2.size(), Rcw, tcw, Ow, intr,
                                             kf_depth_left.data(), kf_bf, xyz.data(), nrm.data(), md.data(), lvl.data(), n_mp, matched_idx.data(),
                                             od.data(), &cnt);
        if (st != MCV_OK) throw std::runtime_error(std::string("FuseMatch: ") + mcv_last_error());
        if (matched_dist) matched_dist->assign(od.begin(), od.begin() + n_mp);
        return (uint)cnt;
    }

    // Object::ComputeBow (src/Object.cpp:238-247): voc.transform(v_desps, bow_vector, bow_feature, 4), once per object
    void ComputeBow(const DBoW3::Vocabulary& voc) {
        if (is_bowed) return;
        std::vector<cv::Mat> v_desps;
        for (int i = 0, sz = desps.rows; i < sz; i++) v_desps.push_back(desps.row(i));
        voc.transform(v_desps, bow_vector, bow_feature, 4);
        is_bowed = true;
    }

    cv::Mat img;
    Keypoints kps;
    Desps desps;
    ORB* extractor = nullptr;
    DBoW3::BowVector bow_vector;
    DBoW3::FeatureVector bow_feature;
    bool is_bowed = false;
    float Rcw[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tcw[3] = {0, 0, 0}, intr[4] = {1, 1, 0, 0};
};

// Tracker::Wnd_Track(obj1, obj2) (src/Tracker.cpp:341-360) for the keypoints of obj1 that own a MapPoint (`mp_kp_idxs`, the
// order of obj1->GetAllMapPointsIdxs()). matched[q] = the obj2 index the reference passes to AddMapPoint, or -1.
inline uint Wnd_Track(const Object& obj1, const std::vector<int>& mp_kp_idxs, const Object& obj2, std::vector<int>& matched,
                      std::vector<int>* matched_dist = nullptr) {
    const int nq = (int)mp_kp_idxs.size();
    matched.assign((size_t)nq, -1);
    std::vector<int32_t> best((size_t)nq + 1), od((size_t)nq + 1);
    int cnt = 0;
    const cv::Mat d1 = (obj1.desps.empty() || obj1.desps.isContinuous()) ? obj1.desps : obj1.desps.clone();
    const cv::Mat d2 = (obj2.desps.empty() || obj2.desps.isContinuous()) ? obj2.desps : obj2.desps.clone();
    const mcv_status st = mcv_wnd_track(mcv_host::kp_ptr(obj1.kps), d1.data, (int)obj1.kps.size(), mp_kp_idxs.data(), nq, mcv_host::kp_ptr(obj2.kps),
                                        d2.data, (int)obj2.kps.size(), obj2.img.cols, obj2.img.rows, matched.data(), best.data(), od.data(), &cnt);
    if (st != MCV_OK) throw std::runtime_error(std::string("Wnd_Track: ") + mcv_last_error());
    if (matched_dist) matched_dist->assign(od.begin(), od.begin() + nq);
    return (uint)cnt;
}
// KL_Track(obj1, obj2, mp2idx) (src/Frame.cpp:34-76). MapPoints are opaque to this layer: mp_kp_idxs[i] =
// obj1->GetMapPointIdx(mps[i]) in the order of obj1->GetMapPointsVector(), seen[i] = mp2idx.count(mps[i]) on entry (updated
// like mp2idx[mps[i]] = ... at :65). Appends the tracked keypoints to obj2.kps exactly like :65-69 (pt = next_pts[i],
// octave = 0), new_idx[i] = the index the reference stores in mp2idx, or -1. Returns cnt.
inline uint KL_Track(const Object& obj1, const std::vector<int>& mp_kp_idxs, Object& obj2, std::vector<uint8_t>& seen, std::vector<int>& new_idx) {
    const int n = (int)mp_kp_idxs.size();
    new_idx.assign((size_t)n, -1);
    seen.resize((size_t)n, 0);
    if (n < 10) return 0;                                                  // :41
    std::vector<cv::KeyPoint> kps((size_t)n), nk((size_t)n);
    for (int i = 0; i < n; ++i) kps[(size_t)i] = obj1.kps[(size_t)mp_kp_idxs[(size_t)i]];
    std::vector<uint8_t> ok((size_t)n);
    int n_ok = 0;
    const cv::Mat a = obj1.img.isContinuous() ? obj1.img : obj1.img.clone(), b = obj2.img.isContinuous() ? obj2.img : obj2.img.clone();
    const mcv_status st = mcv_kl_track(a.data, b.data, a.cols, a.rows, (size_t)a.cols, mcv_host::kp_ptr(kps), n, mcv_host::kp_ptr(nk), ok.data(), &n_ok);
    if (st != MCV_OK) throw std::runtime_error(std::string("KL_Track: ") + mcv_last_error());
    uint cnt = 0;
    for (int i = 0; i < n; ++i) {
        if (!ok[(size_t)i] || seen[(size_t)i]) continue;                   // :57-63
        seen[(size_t)i] = 1;
        new_idx[(size_t)i] = (int)obj2.kps.size();
        obj2.kps.push_back(nk[(size_t)i]);
        cnt++;
    }
    return cnt;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cpp:101-150) on the descriptors the point was observed with
// (`all_ob_desps`, the vector the reference builds at :108-115). Returns what the reference assigns to MapPoint::desp — a
// clone of the row with the least median distance to the others — or an empty Mat where the reference returns early.
inline cv::Mat ComputeDistinctiveDescriptors(const std::vector<cv::Mat>& all_ob_desps, int* best_idx_out = nullptr, int* best_median_out = nullptr) {
    if (best_idx_out) *best_idx_out = -1;
    if (all_ob_desps.empty()) return cv::Mat();
    std::vector<uint8_t> rows((size_t)all_ob_desps.size() * 32);
    for (size_t i = 0; i < all_ob_desps.size(); ++i) memcpy(rows.data() + i * 32, all_ob_desps[i].data, 32);
    const int32_t off[2] = {0, (int32_t)all_ob_desps.size()};
    int32_t bi = -1, bm = -1;
    const mcv_status st = mcv_distinctive_descriptors(rows.data(), off, 1, &bi, &bm, nullptr);
    if (st != MCV_OK) throw std::runtime_error(std::string("ComputeDistinctiveDescriptors: ") + mcv_last_error());
    if (best_idx_out) *best_idx_out = bi;
    if (best_median_out) *best_median_out = bm;
    return all_ob_desps[(size_t)bi].clone();
}
using ObjectRef = std::shared_ptr<Object>;

class Frame {
   public:
    // Frame::Frame stages ORBE + SMatch (src/Frame.cpp:78-138) with an empty optical-flow list (as test/matching_benchmark.cpp:43).
    Frame(const cv::Mat& imgleft, const cv::Mat& imgright, const cv::Mat& imgwide) {
        LEFT = std::make_shared<Object>(imgleft, &extractor_left());
        RIGHT = std::make_shared<Object>(imgright, &extractor_right());
        WIDE = std::make_shared<Object>(imgwide, &extractor_wide());
        // "ORBE": the reference fans the three Extract calls out on ThreadPool(3) (src/Frame.cpp:118-126); each extractor
        // owns a CUDA stream, so three host threads give three concurrent streams.
        auto job = [](ObjectRef o) { return o->extractor->Extract(o->img, o->kps, o->desps); };
        auto l = std::async(std::launch::async, job, LEFT), r = std::async(std::launch::async, job, RIGHT), w = std::async(std::launch::async, job, WIDE);
        const int nl = l.get(), nr = r.get(), nw = w.get();
        if (nl < 0 || nr < 0 || nw < 0) throw std::runtime_error("Frame: extraction failed: " + LEFT->extractor->last_error());
        depth_left.resize(LEFT->size(), -1);
        ComputeStereoMatch(LEFT, RIGHT);  // "SMatch"
    }

    // Frame::ComputeStereoMatch (src/Frame.cpp:150-328): reads both extractors' pyramids (device copies), writes u_right / depth_left.
    void ComputeStereoMatch(ObjectRef left, ObjectRef right) {
        u_right.assign(left->size(), -1.f);
        depth_left.assign(left->size(), -1.f);
        if (left->size() == 0) return;
        const mcv_status st = mcv_stereo_match(left->extractor->handle(), right->extractor->handle(), mcv_host::kp_ptr(left->kps), left->desps.data,
                                               (int)left->size(), mcv_host::kp_ptr(right->kps), right->desps.data, (int)right->size(), bf(), b(),
                                               u_right.data(), depth_left.data(), nullptr, nullptr);
        if (st != MCV_OK) throw std::runtime_error(std::string("ComputeStereoMatch: ") + mcv_last_error());
    }

    // Frame::Parse (src/Frame.cpp:344-364): bf, baseline and the three extractor configs.
    static int Parse(const std::string& config_file) {
        auto root = mcv_host::parse_flat_yaml(config_file);
        bf() = std::stof(root.at("bf"));
        b() = std::stof(root.at("baseline"));
        extractor_left() = ORB(root.at("left_extractor_path"));
        extractor_right() = ORB(root.at("right_extractor_path"));
        extractor_wide() = ORB(root.at("wide_extractor_path"));
        return 0;
    }

    ObjectRef LEFT, RIGHT, WIDE;
    std::vector<float> u_right;
    std::vector<float> depth_left;
    static float& bf() { static float v = 955.40503f; return v; }
    static float& b() { static float v = 1.f; return v; }
    static ORB& extractor_left() { static ORB e; return e; }
    static ORB& extractor_right() { static ORB e; return e; }
    static ORB& extractor_wide() { static ORB e; return e; }
};

// Batched form of the Frame constructor for throughput: n_frames triplets per call through mcv_rig_process (one launch per
// stage over the whole batch, chunks pipelined over three streams). Results are indexed [frame][camera][keypoint].
class Rig {
   public:
    Rig(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, float bf, float baseline, int device = 0) {
        mcv_rig_params p{{nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST}, bf, baseline};
        mcv_rig* raw = nullptr;
        const mcv_status st = mcv_rig_create(&p, device, nullptr, &raw);
        if (st != MCV_OK) throw std::runtime_error(std::string("Rig: ") + mcv_last_error() + " (status " + std::to_string(st) + ")");
        r_.reset(raw, [](mcv_rig* r) { mcv_rig_destroy(r); });
        cap_ = mcv_rig_max_keypoints(raw);
    }
    int cap() const { return cap_; }
    // imgs: [n_frames][3][h][w] u8 host. Outputs sized by the callee: kps/desc [n_frames*3][cap], counts [n_frames*3], u_right/depth [n_frames][cap].
    void Process(const uint8_t* imgs, int n_frames, int w, int h, std::vector<cv::KeyPoint>& kps, std::vector<uint8_t>& desc, std::vector<int32_t>& counts,
                 std::vector<float>& u_right, std::vector<float>& depth_left) {
        cap_ = std::max(cap_, mcv_rig_max_keypoints_for(r_.get(), w, h));   // wide panoramas have more than 4 quadtree roots per level
        kps.resize((size_t)n_frames * 3 * cap_); desc.resize((size_t)n_frames * 3 * cap_ * 32); counts.resize((size_t)n_frames * 3);
        u_right.resize((size_t)n_frames * cap_); depth_left.resize((size_t)n_frames * cap_);
        const mcv_status st = mcv_rig_process(r_.get(), imgs, n_frames, w, h, 0, mcv_host::kp_ptr(kps), desc.data(), counts.data(), u_right.data(),
                                              depth_left.data(), cap_, 0);
        if (st != MCV_OK) throw std::runtime_error(std::string("Rig::Process: ") + mcv_last_error());
    }
    mcv_rig* handle() const { return r_.get(); }

   private:
    std::shared_ptr<mcv_rig> r_;
    int cap_ = 0;
};

}  // namespace MCVSLAM
