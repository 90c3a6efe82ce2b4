// ORACLE — TEST INFRASTRUCTURE ONLY. Not product code (see ora_primitives.hpp header).
//
// CPU restatement of Sologala/MCVSLAM's ORB-extract + Hamming-match hot path, function by function, with the
// reference file:line each block follows. First-party logic (cell grid, quadtree with std::priority_queue,
// IC_Angle, steered rBRIEF, streaming top-2, filters, stereo, projection) is restated over PODs; the OpenCV
// primitives come from ora_primitives.hpp. Build like the reference (CMakeLists.txt:4-6,18-19): -O3, no -march,
// no -ffast-math, libstdc++ (heap / sort tie order is part of the result).
//
// The reference itself cannot be compiled here (needs OpenCV C++, Eigen, Boost, ROS, OSG, pyp — none present and
// no network), so there is no oracle/_ref; see DESIGN.md.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <memory>
#include <queue>
#include <thread>
#include <map>
#include <limits>
#include <vector>

#include "ora_primitives.hpp"
#include "ora_lk.hpp"

using namespace std;  // as the reference does (ORBextractor.cc:67) — this is what makes cos(float) resolve to cosf

namespace ora {

struct KeyPoint {  // layout of cv::KeyPoint (28 bytes)
    float x, y, size, angle, response; int32_t octave, class_id;
};
struct DMatch {  // layout of cv::DMatch (16 bytes)
    int32_t queryIdx, trainIdx, imgIdx; float distance;
};
static_assert(sizeof(KeyPoint) == 28 && sizeof(DMatch) == 16, "layout");

static const int PATCH_SIZE = 31, HALF_PATCH_SIZE = 15, EDGE_THRESHOLD = 19;  // ORBextractor.cc:71-73

static const int8_t BIT_PATTERN_31[256 * 4] = {
#include "ora_pattern.inc"
};

struct OwnedImg {
    vector<uint8_t> buf; int w = 0, h = 0;
    Img view() const { return Img{buf.data(), w, h, (size_t)w}; }
};

// ---------------------------------------------------------------------------------------------------------
// ORBextractor state + init — ORBextractor.cc:402-457
// ---------------------------------------------------------------------------------------------------------
struct Extractor {
    int nfeatures, nlevels, iniThFAST, minThFAST; float scaleFactor;
    vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    vector<int> mnFeaturesPerLevel, umax;
    vector<OwnedImg> mvImagePyramid;  // un-bordered level images (the ROI the reference exposes)
    // debug taps (filled by extract when enabled)
    bool keep_debug = false;
    vector<vector<KeyPoint>> dbg_candidates, dbg_distributed;
    vector<OwnedImg> dbg_blurred;

    void init(int nf, float sf, int nl, int ini, int mn) {
        nfeatures = nf; scaleFactor = sf; nlevels = nl; iniThFAST = ini; minThFAST = mn;
        mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
        mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
        for (int i = 1; i < nlevels; i++) {
            mvScaleFactor[i] = mvScaleFactor[i - 1] * scaleFactor;
            mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
        }
        mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        for (int i = 0; i < nlevels; i++) {
            mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
            mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
        }
        mvImagePyramid.resize(nlevels);
        mnFeaturesPerLevel.resize(nlevels);
        float factor = 1.0f / scaleFactor;
        float nDesiredFeaturesPerScale = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
        int sumFeatures = 0;
        for (int level = 0; level < nlevels - 1; level++) {
            mnFeaturesPerLevel[level] = cv_round(nDesiredFeaturesPerScale);
            sumFeatures += mnFeaturesPerLevel[level];
            nDesiredFeaturesPerScale *= factor;
        }
        mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sumFeatures, 0);
        // umax — ORBextractor.cc:444-456
        umax.assign(HALF_PATCH_SIZE + 1, 0);
        int v, v0, vmax = cv_floor(HALF_PATCH_SIZE * sqrt(2.f) / 2 + 1);
        int vmin = (int)std::ceil(HALF_PATCH_SIZE * sqrt(2.f) / 2);
        const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
        for (v = 0; v <= vmax; ++v) umax[v] = cv_round(sqrt(hp2 - v * v));
        for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }

    // ComputePyramid — ORBextractor.cc:901-919 (level l resized from level l-1; borders never read by this path)
    void ComputePyramid(const Img& image) {
        for (int level = 0; level < nlevels; ++level) {
            float scale = mvInvScaleFactor[level];
            int sw = cv_round((float)image.w * scale), sh = cv_round((float)image.h * scale);
            OwnedImg& L = mvImagePyramid[level];
            L.w = sw; L.h = sh; L.buf.assign((size_t)sw * sh, 0);
            if (level != 0) {
                resize_linear_u8(mvImagePyramid[level - 1].view(), L.buf.data(), sw, sh, sw);
            } else {
                for (int y = 0; y < sh; ++y) memcpy(&L.buf[(size_t)y * sw], image.row(y), sw);
            }
        }
    }
};

// IC_Angle — ORBextractor.cc:75-98
static float IC_Angle(const Img& image, float ptx, float pty, const vector<int>& u_max) {
    int m_01 = 0, m_10 = 0;
    const uint8_t* center = image.row(cv_round(pty)) + cv_round(ptx);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    int step = (int)image.stride;
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        int d = u_max[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return fast_atan2((float)m_01, (float)m_10);
}

// computeOrbDescriptor — ORBextractor.cc:100-141
static const float factorPI = (float)(3.14159265358979323846 / 180.f);
static void computeOrbDescriptor(const KeyPoint& kpt, const Img& img, uint8_t* desc) {
    float angle = (float)kpt.angle * factorPI;
    float a = (float)cos(angle), b = (float)sin(angle);  // float overloads -> glibc cosf/sinf (sincosf)
    const uint8_t* center = img.row(cv_round(kpt.y)) + cv_round(kpt.x);
    const int step = (int)img.stride;
    const int8_t* pattern = BIT_PATTERN_31;
#define ORA_GET_VALUE(idx) \
    center[cv_round((float)pattern[2 * (idx)] * b + (float)pattern[2 * (idx) + 1] * a) * step + \
           cv_round((float)pattern[2 * (idx)] * a - (float)pattern[2 * (idx) + 1] * b)]
    for (int i = 0; i < 32; ++i, pattern += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            int t0 = ORA_GET_VALUE(2 * k), t1 = ORA_GET_VALUE(2 * k + 1);
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
#undef ORA_GET_VALUE
}

// ExtractorNode + DivideNode — ORBextractor.h:33-43, ORBextractor.cc:469-522
struct Pt2i { int x = 0, y = 0; };
struct ExtractorNode;
using ExtractorNodeRef = shared_ptr<ExtractorNode>;
struct ExtractorNode {
    vector<KeyPoint> vKeys; Pt2i UL, UR, BL, BR; vector<ExtractorNodeRef> sons;
    int size() { return (int)vKeys.size(); }
    vector<ExtractorNodeRef>& DivideNode() {
        const int halfX = (int)ceil(static_cast<float>(UR.x - UL.x) / 2);
        const int halfY = (int)ceil(static_cast<float>(BR.y - UL.y) / 2);
        ExtractorNodeRef n1 = make_shared<ExtractorNode>(), n2 = make_shared<ExtractorNode>(),
                         n3 = make_shared<ExtractorNode>(), n4 = make_shared<ExtractorNode>();
        n1->UL = UL; n1->UR = {UL.x + halfX, UL.y}; n1->BL = {UL.x, UL.y + halfY}; n1->BR = {UL.x + halfX, UL.y + halfY};
        n2->UL = n1->UR; n2->UR = UR; n2->BL = n1->BR; n2->BR = {UR.x, UL.y + halfY};
        n3->UL = n1->BL; n3->UR = n1->BR; n3->BL = BL; n3->BR = {n1->BR.x, BL.y};
        n4->UL = n3->UR; n4->UR = n2->BR; n4->BL = n3->BR; n4->BR = BR;
        for (size_t i = 0; i < vKeys.size(); i++) {
            const KeyPoint& kp = vKeys[i];
            if (kp.x < n1->UR.x) {
                if (kp.y < n1->BR.y) n1->vKeys.push_back(kp); else n3->vKeys.push_back(kp);
            } else if (kp.y < n1->BR.y) n2->vKeys.push_back(kp);
            else n4->vKeys.push_back(kp);
        }
        if (n1->size()) sons.push_back(n1);
        if (n2->size()) sons.push_back(n2);
        if (n3->size()) sons.push_back(n3);
        if (n4->size()) sons.push_back(n4);
        return sons;
    }
};

// DistributeOctTree — ORBextractor.cc:524-580 (max-heap keyed on vKeys.size() only; libstdc++ heap order)
static vector<KeyPoint> DistributeOctTree(const vector<KeyPoint>& vToDistributeKeys, const int& minX, const int& maxX,
                                          const int& minY, const int& maxY, const int& N) {
    const int nIni = (int)round(static_cast<float>(maxX - minX) / (maxY - minY));
    const float hX = static_cast<float>(maxX - minX) / nIni;
    auto cmp = [](const ExtractorNodeRef& a, const ExtractorNodeRef& b) { return a->vKeys.size() < b->vKeys.size(); };
    priority_queue<ExtractorNodeRef, vector<ExtractorNodeRef>, decltype(cmp)> q_nodes(cmp);
    vector<ExtractorNodeRef> init_nodes;
    for (int i = 0; i < nIni; i++) {
        ExtractorNodeRef pnode = make_shared<ExtractorNode>();
        pnode->UL = {(int)(hX * static_cast<float>(i)), 0};
        pnode->UR = {(int)(hX * static_cast<float>(i + 1)), 0};
        pnode->BL = {pnode->UL.x, maxY - minY};
        pnode->BR = {pnode->UR.x, maxY - minY};
        init_nodes.push_back(pnode);
    }
    for (size_t i = 0, sz = vToDistributeKeys.size(); i < sz; i++) {
        const KeyPoint& kp = vToDistributeKeys[i];
        init_nodes[(size_t)(kp.x / hX)]->vKeys.push_back(kp);
    }
    for (unsigned i = 0, sz = (unsigned)init_nodes.size(); i < sz; i++) {
        ExtractorNodeRef pnode = init_nodes[i];
        if (pnode->size() == 0) continue;
        q_nodes.push(pnode);
    }
    vector<KeyPoint> ret;
    if (q_nodes.empty()) return ret;  // guard: reference would call top() on an empty heap when N > 0
    while (q_nodes.size() < (size_t)N) {
        auto pnode = q_nodes.top();
        if (pnode->size() == 1) break;
        q_nodes.pop();
        for (auto& node : pnode->DivideNode()) q_nodes.push(node);
    }
    ret.reserve(q_nodes.size());
    while (false == q_nodes.empty()) {
        auto pnode = q_nodes.top();
        q_nodes.pop();
        KeyPoint max_response_kp = pnode->vKeys[0];
        for (unsigned i = 1, sz = (unsigned)pnode->size(); i < sz; i++)
            if (pnode->vKeys[i].response > max_response_kp.response) max_response_kp = pnode->vKeys[i];
        ret.push_back(max_response_kp);
    }
    return ret;
}

// ComputeKeyPointsOctTree — ORBextractor.cc:582-654
static int ComputeKeyPointsOctTree(Extractor& E, vector<vector<KeyPoint>>& allKeypoints) {
    allKeypoints.assign(E.nlevels, {});
    if (E.keep_debug) { E.dbg_candidates.assign(E.nlevels, {}); E.dbg_distributed.assign(E.nlevels, {}); }
    const float W = 35;
    vector<FastKp> cell;
    for (int level = 0; level < E.nlevels; ++level) {
        const Img im = E.mvImagePyramid[level].view();
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
        const int maxBorderX = im.w - EDGE_THRESHOLD + 3, maxBorderY = im.h - EDGE_THRESHOLD + 3;
        vector<KeyPoint> vToDistributeKeys;
        vToDistributeKeys.reserve(E.nfeatures * 10);
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols < 1 || nRows < 1) return -2;  // guard: reference divides by zero here
        const int wCell = (int)ceil(width / nCols), hCell = (int)ceil(height / nRows);
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                Img roi{im.row((int)iniY) + (int)iniX, (int)maxX - (int)iniX, (int)maxY - (int)iniY, im.stride};
                fast9_16_nms(roi, E.iniThFAST, cell);
                if (cell.empty()) fast9_16_nms(roi, E.minThFAST, cell);
                for (const FastKp& k : cell) {
                    KeyPoint kp{(float)k.x, (float)k.y, 7.f, -1.f, (float)k.score, 0, -1};
                    kp.x += j * wCell;
                    kp.y += i * hCell;
                    vToDistributeKeys.push_back(kp);
                }
            }
        }
        vector<KeyPoint>& keypoints = allKeypoints[level];
        if ((maxBorderY - minBorderY) <= 0 || (int)round(width / (maxBorderY - minBorderY)) < 1) return -2;
        keypoints = DistributeOctTree(vToDistributeKeys, minBorderX, maxBorderX, minBorderY, maxBorderY, E.mnFeaturesPerLevel[level]);
        if (E.keep_debug) { E.dbg_candidates[level] = vToDistributeKeys; E.dbg_distributed[level] = keypoints; }
        const int scaledPatchSize = (int)(PATCH_SIZE * E.mvScaleFactor[level]);
        for (size_t i = 0; i < keypoints.size(); i++) {
            keypoints[i].x += minBorderX;
            keypoints[i].y += minBorderY;
            keypoints[i].octave = level;
            keypoints[i].size = (float)scaledPatchSize;
        }
    }
    for (int level = 0; level < E.nlevels; ++level)
        for (KeyPoint& kp : allKeypoints[level]) kp.angle = IC_Angle(E.mvImagePyramid[level].view(), kp.x, kp.y, E.umax);
    return 0;
}

// ORBextractor::operator() — ORBextractor.cc:831-899. `kps` is in/out (pre-seeded keypoints are appended per octave).
static int Extract(Extractor& E, const Img& image, vector<KeyPoint>& kps, vector<uint8_t>& desc) {
    if (image.data == nullptr || image.w <= 0 || image.h <= 0) return -1;
    E.ComputePyramid(image);
    vector<vector<KeyPoint>> allKeypoints;
    int rc = ComputeKeyPointsOctTree(E, allKeypoints);
    if (rc) return rc;
    for (const KeyPoint& kp : kps) {
        if (kp.octave < 0 || kp.octave >= E.nlevels) return -3;  // guard: reference indexes out of range
        allKeypoints[kp.octave].push_back(kp);
    }
    int nkp = 0;
    for (int level = 0; level < E.nlevels; ++level) nkp += (int)allKeypoints[level].size();
    desc.assign((size_t)nkp * 32, 0);
    kps.assign(nkp, KeyPoint{});
    if (E.keep_debug) E.dbg_blurred.assign(E.nlevels, {});
    int offset = 0, cnt = 0;
    for (int level = 0; level < E.nlevels; ++level) {
        vector<KeyPoint>& keypoints = allKeypoints[level];
        int nkeypointsLevel = (int)keypoints.size();
        if (nkeypointsLevel == 0) continue;
        const OwnedImg& L = E.mvImagePyramid[level];
        OwnedImg work; work.w = L.w; work.h = L.h; work.buf.resize(L.buf.size());
        gauss7_u8(L.view(), work.buf.data(), L.w);
        for (int i = 0; i < nkeypointsLevel; ++i) computeOrbDescriptor(keypoints[i], work.view(), &desc[(size_t)(offset + i) * 32]);
        if (E.keep_debug) E.dbg_blurred[level] = work;
        offset += nkeypointsLevel;
        float scale = E.mvScaleFactor[level];
        for (KeyPoint& kp : keypoints) {
            if (level != 0) { kp.x *= scale; kp.y *= scale; }
            kps[cnt++] = kp;
        }
    }
    return cnt;
}

// ---------------------------------------------------------------------------------------------------------
// Matcher — src/Matcher.cpp
// ---------------------------------------------------------------------------------------------------------
// LoopBody / KnnMatch(vector<Mat>, vector<Mat>) — src/Matcher.cpp:245-302: streaming top-2, sentinel (0, 999).
static void knn2_stream(const uint8_t* q, const uint8_t* const* cand, int ncand, unsigned d[2], unsigned d_idx[2]) {
    d[0] = d[1] = 999; d_idx[0] = d_idx[1] = 0;
    for (int j = 0; j < ncand; j++) {
        unsigned dist = hamming256(q, cand[j]);
        if (dist < d[0]) { d[1] = d[0]; d_idx[1] = d_idx[0]; d[0] = dist; d_idx[0] = j; }
        else if (dist < d[1]) { d[1] = dist; d_idx[1] = j; }
    }
}

// MatchResKnn::FilterRatio on one 2-entry row — src/Matcher.cpp:100-111
static inline bool ratio_pass(float d0, float d1, float ratio) { return d0 / d1 <= ratio; }

// ---------------------------------------------------------------------------------------------------------
// Frame::ComputeStereoMatch — src/Frame.cpp:150-328
// ---------------------------------------------------------------------------------------------------------
static int ComputeStereoMatch(const Extractor& EL, const Extractor& ER, const KeyPoint* kl, const uint8_t* dl, int nl,
                              const KeyPoint* kr, const uint8_t* dr, int nr, int nRows, float bf, float b,
                              float* u_right, float* depth_left, int* best_dist_out, int* best_r_out) {
    for (int i = 0; i < nl; ++i) { u_right[i] = -1.0f; depth_left[i] = -1.0f; if (best_dist_out) best_dist_out[i] = -1; if (best_r_out) best_r_out[i] = -1; }
    vector<vector<size_t>> vRowIndices(nRows);
    for (int iR = 0; iR < nr; iR++) {
        const KeyPoint& kp = kr[iR];
        const float kpY = kp.y;
        const float r = 10.f * EL.mvScaleFactor[kp.octave];
        const int maxr = (int)ceil(kpY + r);
        const int minr = (int)floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++) {
            if (yi < 0 || yi >= nRows) continue;  // guard: reference indexes unchecked (UB); unreachable for quadtree kps
            vRowIndices[yi].push_back(iR);
        }
    }
    const float minZ = b, minD = 1;
    float maxD = min(bf / minZ, float(1000.0));
    vector<pair<int, int>> vDistIdx;
    for (int iL = 0; iL < nl; iL++) {
        const KeyPoint& kpL = kl[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        if (vL < 0 || (size_t)vL >= (size_t)nRows) continue;  // guard
        const vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
        if (vCandidates.size() == 0) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        vector<const uint8_t*> right_desps; vector<unsigned> ori_idx_right;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            const KeyPoint& kpR = kr[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) { right_desps.push_back(dr + iR * 32); ori_idx_right.push_back((unsigned)iR); }
        }
        if (right_desps.size() == 0) continue;
        unsigned d[2], di[2];
        knn2_stream(dl + (size_t)iL * 32, right_desps.data(), (int)right_desps.size(), d, di);
        // FilterRatio(0.70) then FilterThreshold(int(46*0.75)=34) — src/Frame.cpp:225
        if (!ratio_pass((float)d[0], (float)d[1], 0.70f)) continue;
        if ((float)d[0] > (int)(46 * 0.75)) continue;
        const unsigned bestIdxR = di[0];
        const float uR0 = kr[ori_idx_right[bestIdxR]].x;
        const float scaleFactor = EL.mvInvScaleFactor[kpL.octave];
        const float scaleduL = round(kpL.x * scaleFactor), scaledvL = round(kpL.y * scaleFactor), scaleduR0 = round(uR0 * scaleFactor);
        const int w = 5;
        const Img IL = EL.mvImagePyramid[kpL.octave].view();
        const Img IRimg = ER.mvImagePyramid[kpL.octave].view();
        if (scaledvL - w < 0 || scaledvL + w + 1 >= IL.h || scaleduL - w < 0 || scaleduL + w + 1 >= IL.w) continue;
        int bestDist = INT_MAX, bestincR = 0;
        const int L = 5;
        float vDists[2 * L + 1];
        const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
        if (iniu < 0 || endu >= IRimg.w) continue;
        if (scaleduR0 - L - w < 0) continue;  // guard: reference would throw on a negative colRange
        const int cyL = (int)scaledvL, cxL = (int)scaleduL;
        const int lc = IL.row(cyL)[cxL];
        for (int incR = -L; incR <= +L; incR++) {
            const int cxR = (int)scaleduR0 + incR;
            const int rc = IRimg.row(cyL)[cxR];
            int sad = 0;  // cv::norm(IL - IL(w,w), IR - IR(w,w), NORM_L1) on CV_16S
            for (int yy = -w; yy <= w; ++yy)
                for (int xx = -w; xx <= w; ++xx)
                    sad += abs((IL.row(cyL + yy)[cxL + xx] - lc) - (IRimg.row(cyL + yy)[cxR + xx] - rc));
            float dist = (float)sad;
            if (dist < bestDist) { bestDist = (int)dist; bestincR = incR; }
            vDists[L + incR] = dist;
        }
        if (bestincR == -L || bestincR == L) continue;
        const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
        const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
        if (deltaR < -1 || deltaR > 1) continue;  // NaN (0/0) falls through both tests, as in the reference
        float bestuR = EL.mvScaleFactor[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
        float disparity = (uL - bestuR);
        if (disparity >= minD && disparity < maxD) {
            depth_left[iL] = bf / disparity;
            u_right[iL] = bestuR;
            if (best_dist_out) best_dist_out[iL] = bestDist;
            if (best_r_out) best_r_out[iL] = (int)ori_idx_right[bestIdxR];
            vDistIdx.push_back(pair<int, int>(bestDist, iL));
        }
    }
    if (vDistIdx.empty()) return 0;  // guard: reference indexes an empty vector
    sort(vDistIdx.begin(), vDistIdx.end());
    const float median = (float)vDistIdx[(size_t)(vDistIdx.size() * 1.0 / 2)].first;
    const float th_max_dist = 1.6f * median;
    const float th_min_dist = (float)(0.4 * median);
    int l = 0, r = (int)vDistIdx.size() - 1;
    for (; r >= 0 && vDistIdx[r].first > th_max_dist; r--) { u_right[vDistIdx[r].second] = -1; depth_left[vDistIdx[r].second] = -1; }
    for (; l < r && vDistIdx[l].first < th_min_dist; l++) { u_right[vDistIdx[l].second] = -1; depth_left[vDistIdx[l].second] = -1; }
    return (int)vDistIdx.size();
}

// ---------------------------------------------------------------------------------------------------------
// Object grid + ProjectBunchMapPoints — src/Object.cpp:182-308, include/Object.hpp:27-32
// ---------------------------------------------------------------------------------------------------------
static const int GRID = 30;
struct FeatureGrid {
    vector<size_t> cell[GRID][GRID]; float winv, hinv; int W, H;
    // AssignFeaturesToGrid / PosInGrid — src/Object.cpp:182-201,249-257 (note: round(), and index 30 is dropped)
    void build(const KeyPoint* kps, int n, int w, int h) {
        W = w; H = h; winv = (float)(static_cast<double>(GRID) / w); hinv = (float)(static_cast<double>(GRID) / h);  // src/Object.cpp:26-27
        for (int i = 0; i < n; ++i) {
            int px = (int)round((kps[i].x - 0) * winv), py = (int)round((kps[i].y - 0) * hinv);
            if (px < 0 || px >= GRID || py < 0 || py >= GRID) continue;
            cell[px][py].push_back(i);
        }
    }
    // GetFeaturesInArea — src/Object.cpp:263-308 (no level gate: minLevel=-1,maxLevel=-1 defaults)
    void query(const KeyPoint* kps, float x, float y, float r, vector<size_t>& out) const {
        out.clear();
        if (x < 0 || y < 0 || x >= W || y >= H) return;
        const int nMinCellX = max(0, (int)floor((x - r) * winv));
        if (nMinCellX >= GRID) return;
        const int nMaxCellX = min(GRID - 1, (int)ceil((x + r) * winv));
        if (nMaxCellX < 0) return;
        const int nMinCellY = max(0, (int)floor((y - r) * hinv));
        if (nMinCellY >= GRID) return;
        const int nMaxCellY = min(GRID - 1, (int)ceil((y + r) * hinv));
        if (nMaxCellY < 0) return;
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                for (size_t j : cell[ix][iy]) {
                    const float distx = kps[j].x - x, disty = kps[j].y - y;
                    if (fabs(distx) < r && fabs(disty) < r) out.push_back(j);
                }
    }
};

}  // namespace ora

// =========================================================================================================
// C interface for ctypes (tests/, bench.py cpu_baseline). All buffers are caller-owned.
// =========================================================================================================
using namespace ora;

extern "C" {

void ora_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride) {
    resize_linear_u8(Img{src, sw, sh, (size_t)sstride}, dst, dw, dh, dstride);
}
// cv::cvtColor(COLOR_BGR2GRAY) on CV_8UC3 as called by System::Track (src/System.cpp:60-64). OpenCV's 8-bit path is fixed
// point with 15 fractional bits: B2Y = 3735, G2Y = 19235, R2Y = 9798, rounded (pinned against cv2 4.13: tests/golden/bgr_golden.npz).
void ora_bgr2gray_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * sstride;
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) d[x] = (uint8_t)((s[3 * x] * 3735 + s[3 * x + 1] * 19235 + s[3 * x + 2] * 9798 + (1 << 14)) >> 15);
    }
}
void ora_gauss7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride) {
    gauss7_u8(Img{src, w, h, (size_t)sstride}, dst, dstride);
}
float ora_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void ora_fast_atan2_array(const float* y, const float* x, float* out, int n) { for (int i = 0; i < n; ++i) out[i] = fast_atan2(y[i], x[i]); }
void ora_sincosf_array(const float* a, float* s, float* c, int n) { for (int i = 0; i < n; ++i) { s[i] = sin(a[i]); c[i] = cos(a[i]); } }
// cv::FAST on a (sub-)image; out = (x, y, score) int triples; returns count (<= cap written)
int ora_fast(const uint8_t* img, int w, int h, int stride, int threshold, int* out, int cap) {
    vector<FastKp> k; fast9_16_nms(Img{img, w, h, (size_t)stride}, threshold, k);
    for (int i = 0; i < (int)k.size() && i < cap; ++i) { out[3 * i] = k[i].x; out[3 * i + 1] = k[i].y; out[3 * i + 2] = k[i].score; }
    return (int)k.size();
}
int ora_hamming256(const uint8_t* a, const uint8_t* b) { return (int)hamming256(a, b); }

// ---- extractor handle ----
void* ora_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    Extractor* E = new Extractor(); E->init(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST); return E;
}
void ora_orb_destroy(void* h) { delete (Extractor*)h; }
void ora_orb_set_debug(void* h, int on) { ((Extractor*)h)->keep_debug = on != 0; }
void ora_orb_params(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int* quota, int* umax16) {
    Extractor* E = (Extractor*)h;
    for (int i = 0; i < E->nlevels; ++i) { scale[i] = E->mvScaleFactor[i]; inv_scale[i] = E->mvInvScaleFactor[i]; sigma2[i] = E->mvLevelSigma2[i]; inv_sigma2[i] = E->mvInvLevelSigma2[i]; quota[i] = E->mnFeaturesPerLevel[i]; }
    for (int i = 0; i < 16; ++i) umax16[i] = E->umax[i];
}
// kps: in/out buffer with n_seeds valid entries on entry and capacity cap; desc: cap*32 bytes. Returns count, or <0.
int ora_orb_extract(void* h, const uint8_t* img, int w, int hgt, int stride, KeyPoint* kps, int n_seeds, uint8_t* desc, int cap) {
    Extractor* E = (Extractor*)h;
    vector<KeyPoint> v(kps, kps + n_seeds); vector<uint8_t> d;
    int n = Extract(*E, Img{img, w, hgt, (size_t)stride}, v, d);
    if (n < 0) return n;
    if (n > cap) return -4;
    memcpy(kps, v.data(), (size_t)n * sizeof(KeyPoint)); memcpy(desc, d.data(), (size_t)n * 32);
    return n;
}
int ora_orb_level_size(void* h, int level, int* w, int* hgt) { Extractor* E = (Extractor*)h; *w = E->mvImagePyramid[level].w; *hgt = E->mvImagePyramid[level].h; return 0; }
void ora_orb_level_copy(void* h, int level, uint8_t* dst) { Extractor* E = (Extractor*)h; memcpy(dst, E->mvImagePyramid[level].buf.data(), E->mvImagePyramid[level].buf.size()); }
void ora_orb_blurred_copy(void* h, int level, uint8_t* dst) { Extractor* E = (Extractor*)h; if (level < (int)E->dbg_blurred.size()) memcpy(dst, E->dbg_blurred[level].buf.data(), E->dbg_blurred[level].buf.size()); }
// which: 0 = FAST candidates (cell order, coords relative to the 16-px border), 1 = quadtree output (heap-pop order)
int ora_orb_debug_kps(void* h, int which, int level, KeyPoint* out, int cap) {
    Extractor* E = (Extractor*)h;
    const vector<KeyPoint>& v = which == 0 ? E->dbg_candidates[level] : E->dbg_distributed[level];
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}
// Static ORBextractor::DistributeOctTree (ORBextractor.h:73-74)
int ora_distribute_octree(const KeyPoint* in, int n, int minX, int maxX, int minY, int maxY, int N, KeyPoint* out, int cap) {
    vector<KeyPoint> v(in, in + n);
    vector<KeyPoint> r = DistributeOctTree(v, minX, maxX, minY, maxY, N);
    for (int i = 0; i < (int)r.size() && i < cap; ++i) out[i] = r[i];
    return (int)r.size();
}

// ---- matcher ----
// KnnMatch(vector<Mat>, vector<Mat>, 2) — always 2 entries/query, sentinel (trainIdx 0, distance 999), imgIdx -1.
void ora_knn2_firstparty(const uint8_t* q, int nq, const uint8_t* t, int nt, DMatch* out /* nq*2 */) {
    vector<const uint8_t*> c(nt);
    for (int j = 0; j < nt; ++j) c[j] = t + (size_t)j * 32;
    for (int i = 0; i < nq; ++i) {
        unsigned d[2], di[2];
        knn2_stream(q + (size_t)i * 32, c.data(), nt, d, di);
        out[2 * i] = DMatch{i, (int)di[0], -1, (float)d[0]};
        out[2 * i + 1] = DMatch{i, (int)di[1], -1, (float)d[1]};
    }
}
// KnnMatch(Mat, Mat, 2) == cv::BFMatcher(NORM_HAMMING).knnMatch: min(2, nt) entries/query sorted by (distance, trainIdx),
// imgIdx 0. out has nq*2 slots; unused slots get trainIdx -1. Returns entries per query.
int ora_knn2_bf(const uint8_t* q, int nq, const uint8_t* t, int nt, DMatch* out) {
    int k = nt < 2 ? nt : 2;
    for (int i = 0; i < nq; ++i) {
        int bd[2] = {INT_MAX, INT_MAX}, bi[2] = {-1, -1};
        for (int j = 0; j < nt; ++j) {
            int d = (int)hamming256(q + (size_t)i * 32, t + (size_t)j * 32);
            if (d < bd[0]) { bd[1] = bd[0]; bi[1] = bi[0]; bd[0] = d; bi[0] = j; }
            else if (d < bd[1]) { bd[1] = d; bi[1] = j; }
        }
        for (int e = 0; e < 2; ++e) out[2 * i + e] = e < k ? DMatch{i, bi[e], 0, (float)bd[e]} : DMatch{i, -1, 0, 0.f};
    }
    return k;
}
// Candidate-list 2-NN (stereo / projection / fuse / BoW call sites): for query i the candidates are
// cand_idx[cand_off[i] .. cand_off[i+1]) (indices into t, in list order). trainIdx = position in the list.
void ora_knn2_candidates(const uint8_t* q, int nq, const uint8_t* t, const int* cand_off, const int* cand_idx, DMatch* out) {
    for (int i = 0; i < nq; ++i) {
        int n = cand_off[i + 1] - cand_off[i];
        vector<const uint8_t*> c(n);
        for (int j = 0; j < n; ++j) c[j] = t + (size_t)cand_idx[cand_off[i] + j] * 32;
        unsigned d[2], di[2];
        knn2_stream(q + (size_t)i * 32, c.data(), n, d, di);
        out[2 * i] = DMatch{i, (int)di[0], -1, (float)d[0]};
        out[2 * i + 1] = DMatch{i, (int)di[1], -1, (float)d[1]};
    }
}
// MatchResKnn::FilterRatio — src/Matcher.cpp:100-111. knn: nq rows of `per` (1 or 2) entries. Returns count.
int ora_filter_ratio(const DMatch* knn, int nq, int per, float ratio, DMatch* out) {
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        if (per == 1) out[n++] = knn[i];
        else if (per >= 2 && knn[per * i].distance / knn[per * i + 1].distance <= ratio) out[n++] = knn[per * i];
    }
    return n;
}
// MatchRes::FilterThreshold — src/Matcher.cpp:23-35 (swap-remove, reorders). In place; returns new size.
int ora_filter_threshold(DMatch* m, int n, int thres_hold) {
    int i = 0, j = n - 1;
    while (i <= j) { if (m[i].distance > thres_hold) m[i] = m[j--]; else i++; }
    return i;
}
// MatchRes::FilterOrientation — src/Matcher.cpp:44-74 (std::sort of 40 bins by size, keep 3 largest).
int ora_filter_orientation(DMatch* m, int n, const KeyPoint* kps1, const KeyPoint* kps2) {
    const int HISTO_LENGTH = 40; const float HISTO_FACTOR = 1.0f / (360.0f / HISTO_LENGTH);
    std::vector<unsigned> rot_bins[40];
    for (int i = 0; i < HISTO_LENGTH; i++) rot_bins[i].reserve(500);
    for (unsigned i = 0, sz = (unsigned)n; i < sz; i++) {
        float rot = kps1[m[i].queryIdx].angle - kps2[m[i].trainIdx].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin_id = (int)round(rot * HISTO_FACTOR);
        if (bin_id == HISTO_LENGTH) bin_id = 0;
        rot_bins[bin_id].push_back(i);
    }
    auto cmp = [](const vector<unsigned>& a, const vector<unsigned>& b) { return a.size() > b.size(); };
    sort(&rot_bins[0], &rot_bins[0] + HISTO_LENGTH, cmp);
    vector<DMatch> ret;
    for (unsigned i = 0; i < 3; i++) for (auto idx : rot_bins[i]) ret.push_back(m[idx]);
    for (size_t i = 0; i < ret.size(); ++i) m[i] = ret[i];
    return (int)ret.size();
}

// MatchRes::FilterFMatrix + CheckDistEpipolarLine — src/Matcher.cpp:76-91,310-325. F12 = 3x3 row-major float (F12.at<float>(r, c)),
// LevelSigma2 = the SECOND object's extractor->mvLevelSigma2 (src/Map.cpp:307). Epipolar line in image 2: l = x1' F12 = [a b c]
// (column sums with kp1), float arithmetic in source order; `den == 0` rejects; the final compare promotes to double
// (3.84 is a double literal). Swap-remove like FilterThreshold: survivors are reordered. In place; returns the new size.
static bool CheckDistEpipolarLine(const KeyPoint& kp1, const KeyPoint& kp2, const float* F12, const float* LevelSigma2) {
    const float a = kp1.x * F12[0 * 3 + 0] + kp1.y * F12[1 * 3 + 0] + F12[2 * 3 + 0];
    const float b = kp1.x * F12[0 * 3 + 1] + kp1.y * F12[1 * 3 + 1] + F12[2 * 3 + 1];
    const float c = kp1.x * F12[0 * 3 + 2] + kp1.y * F12[1 * 3 + 2] + F12[2 * 3 + 2];
    const float num = a * kp2.x + b * kp2.y + c;
    const float den = a * a + b * b;
    if (den == 0) return false;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * LevelSigma2[kp2.octave];
}
int ora_filter_fmatrix(DMatch* m, int n, const KeyPoint* kps1, const KeyPoint* kps2, const float* F12, const float* level_sigma2) {
    int i = 0, j = n - 1;
    while (i <= j) {
        DMatch d = m[i];
        if (false == CheckDistEpipolarLine(kps1[d.queryIdx], kps2[d.trainIdx], F12, level_sigma2)) m[i] = m[j--]; else i++;
    }
    return i;
}

// ---- stereo (uses the pyramids held by the two extractor handles, as the reference does) ----
int ora_stereo_match(void* hl, void* hr, const KeyPoint* kl, const uint8_t* dl, int nl, const KeyPoint* kr, const uint8_t* dr, int nr,
                     int nRows, float bf, float b, float* u_right, float* depth_left, int* best_dist, int* best_r) {
    return ComputeStereoMatch(*(Extractor*)hl, *(Extractor*)hr, kl, dl, nl, kr, dr, nr, nRows, bf, b, u_right, depth_left, best_dist, best_r);
}

// ---- projection (ProjectBunchMapPoints over an ORDERED MapPoint array; set iteration order is the caller's) ----
// Rcw row-major 3x3, tcw 3, intr = fx fy cx cy. Per MapPoint: out_idx = matched keypoint index or -1, out_dist = best
// Hamming distance (or -1). Returns the number of successful projections (`cnt`).
int ora_project_match(const KeyPoint* kps, const uint8_t* desc, int n, int W, int H, const float* scale_factors,
                      const float* Rcw, const float* tcw, const float* intr, const float* mp_xyz, const uint8_t* mp_desc,
                      const int* mp_level, int n_mp, float r_threshold, int* out_idx, int* out_dist) {
    FeatureGrid* G = new FeatureGrid();
    G->build(kps, n, W, H);
    int cnt = 0; vector<size_t> ori; vector<const uint8_t*> c;
    for (int m = 0; m < n_mp; ++m) {
        out_idx[m] = -1; out_dist[m] = -1;
        const float* P = mp_xyz + 3 * m;
        // cv::Mat float GEMM: Pc = R*Pw + t (accumulates left to right in double inside cv::gemm? no: float)
        float pc[3];
        for (int r = 0; r < 3; ++r) pc[r] = (Rcw[3 * r] * P[0] + Rcw[3 * r + 1] * P[1] + Rcw[3 * r + 2] * P[2]) + tcw[r];
        if (pc[2] < 0) continue;
        // Pinhole::project — modules/camera/Pinhole.cpp:45-47
        const float u = intr[0] * pc[0] / pc[2] + intr[2], v = intr[1] * pc[1] / pc[2] + intr[3];
        const float r = r_threshold * scale_factors[mp_level[m]];
        G->query(kps, u, v, r, ori);
        if (ori.empty()) continue;
        c.resize(ori.size());
        for (size_t j = 0; j < ori.size(); ++j) c[j] = desc + ori[j] * 32;
        unsigned d[2], di[2];
        knn2_stream(mp_desc + (size_t)m * 32, c.data(), (int)c.size(), d, di);
        if (!ratio_pass((float)d[0], (float)d[1], 0.6f)) continue;   // FilterRatio() default 0.6
        if ((float)d[0] > 46) continue;                              // FilterThreshold() default ORB_GOOD_THRESHOLD
        out_idx[m] = (int)ori[di[0]]; out_dist[m] = (int)d[0]; cnt++;
    }
    delete G;
    return cnt;
}

// ---- Map::Fuse matching front-end (src/Map.cpp:478-527) over an ORDERED MapPoint array ----
// Per MapPoint: viewing-angle gate (:487-488), Object::Map + z gate (:490-492), projection (:494), GetFeaturesInArea(uv, 10)
// (:495), per candidate the level gate (:503) and the reprojection gates (:505-524 — as written: the stereo branch compares the
// predicted right coordinate with kf->depth_left[idx] and scales by mvLevelSigma2, the mono branch by mvInvLevelSigma2), then
// KnnMatch({desp}, desps).FilterRatio().FilterThreshold() (:527). out_idx[m] = ori_kp_idx[res[0].trainIdx] or -1. The map
// surgery that follows (:528-547) is pointer-graph state and stays with the caller.
// cv::Mat arithmetic restated from OpenCV: Mat - Mat and the 3x3 gemm in float (pinned against cv2.gemm), cv::norm and
// Mat::dot accumulate the float products in double.
int ora_fuse_match(const KeyPoint* kps, const uint8_t* desc, int n, int W, int H, const float* sigma2, const float* inv_sigma2,
                   const float* Rcw, const float* tcw, const float* Ow, const float* intr, const float* depth_left, float bf,
                   const float* mp_xyz, const float* mp_normal, const uint8_t* mp_desc, const int* mp_level, int n_mp, int* out_idx,
                   int* out_dist) {
    FeatureGrid* G = new FeatureGrid();
    G->build(kps, n, W, H);
    int cnt = 0; vector<size_t> cand, ori; vector<const uint8_t*> c;
    for (int m = 0; m < n_mp; ++m) {
        out_idx[m] = -1; out_dist[m] = -1;
        const float* P = mp_xyz + 3 * m; const float* Pn = mp_normal + 3 * m;
        const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
        const float dist3D = (float)sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);   // cv::norm
        const double dot = (double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2];                   // Mat::dot
        if (dot < 0.5 * dist3D) continue;
        float pc[3];
        for (int r = 0; r < 3; ++r) pc[r] = (Rcw[3 * r] * P[0] + Rcw[3 * r + 1] * P[1] + Rcw[3 * r + 2] * P[2]) + tcw[r];
        const float z = pc[2];
        if (z <= 0) continue;
        const float invz = 1. / z;
        const float u = intr[0] * pc[0] / pc[2] + intr[2], v = intr[1] * pc[1] / pc[2] + intr[3];
        G->query(kps, u, v, 10, cand);
        ori.clear();
        for (size_t idx : cand) {
            const KeyPoint& kp = kps[idx];
            const int level = kp.octave;
            if (level < max(0, int(mp_level[m] - 1)) || level > mp_level[m]) continue;
            if (depth_left[idx] >= 0) {
                const float ur = u - bf * invz;
                const float ex = u - kp.x, ey = v - kp.y, er = ur - depth_left[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * sigma2[level] > 7.8) continue;
            } else {
                const float ex = u - kp.x, ey = v - kp.y;
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv_sigma2[level] > 5.99) continue;
            }
            ori.push_back(idx);
        }
        // KnnMatch({desp}, {}) leaves the sentinel pair (distance 999, 999) and FilterRatio drops it (999 / 999 > 0.6): an empty
        // candidate list never matches (src/Matcher.cpp:256-275,100-111)
        if (ori.empty()) continue;
        c.resize(ori.size());
        for (size_t j = 0; j < ori.size(); ++j) c[j] = desc + ori[j] * 32;
        unsigned d[2], di[2];
        knn2_stream(mp_desc + (size_t)m * 32, c.data(), (int)c.size(), d, di);
        if (!ratio_pass((float)d[0], (float)d[1], 0.6f)) continue;
        if ((float)d[0] > 46) continue;
        out_idx[m] = (int)ori[di[0]]; out_dist[m] = (int)d[0]; cnt++;
    }
    delete G;
    return cnt;
}

// ---- Tracker::Wnd_Track (src/Tracker.cpp:341-360): every keypoint of obj1 that owns a MapPoint (q_idx, in the caller's order)
// is matched inside a +-20 px window of obj2. As written, the reference takes candi_idxs[t_res[0].queryIdx] — queryIdx is
// always 0 — so a successful match reports the FIRST candidate of the window, not the best one; out_idx reproduces that, and
// out_best holds the index the match really belongs to. FilterOrientation on a single match keeps it (its bin is the fullest).
int ora_wnd_track(const KeyPoint* kps1, const uint8_t* desc1, const int* q_idx, int n_q, const KeyPoint* kps2, const uint8_t* desc2, int n2,
                  int W, int H, int* out_idx, int* out_best, int* out_dist) {
    FeatureGrid* G = new FeatureGrid();
    G->build(kps2, n2, W, H);
    int cnt = 0; vector<size_t> cand; vector<const uint8_t*> c;
    for (int q = 0; q < n_q; ++q) {
        out_idx[q] = -1; out_best[q] = -1; out_dist[q] = -1;
        const KeyPoint& kp = kps1[q_idx[q]];
        G->query(kps2, kp.x, kp.y, 20, cand);
        if (cand.empty()) continue;
        c.resize(cand.size());
        for (size_t j = 0; j < cand.size(); ++j) c[j] = desc2 + cand[j] * 32;
        unsigned d[2], di[2];
        knn2_stream(desc1 + (size_t)q_idx[q] * 32, c.data(), (int)c.size(), d, di);
        if (!ratio_pass((float)d[0], (float)d[1], 0.6f)) continue;
        if ((float)d[0] > 46) continue;
        out_idx[q] = (int)cand[0]; out_best[q] = (int)cand[di[0]]; out_dist[q] = (int)d[0]; cnt++;
    }
    delete G;
    return cnt;
}

// ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cpp:101-150), one call per MapPoint in a batch. desc rows
// [off[m], off[m + 1]) = all_ob_desps of point m (:108-115). Literal restatement: N x N int buffer with a zero diagonal (:120-128),
// std::sort of every row, median = row[0.5 * (N - 1)] (double product converted to the index type), strict '<' against INT_MAX
// (:131-141). best_idx = -1 where the reference returns early (:104,117).
void ora_distinctive(const uint8_t* desc, const int* off, int n_mp, int* best_idx, int* best_median) {
    for (int m = 0; m < n_mp; ++m) {
        const uint8_t* d = desc + (size_t)off[m] * 32;
        const size_t N = (size_t)(off[m + 1] - off[m]);
        best_idx[m] = -1; best_median[m] = -1;
        if (N == 0) continue;
        vector<vector<int>> buf(N, vector<int>(N, 0));
        for (size_t i = 0; i < N; i++)
            for (size_t j = i + 1; j < N; j++) {
                const int distij = (int)hamming256(d + i * 32, d + j * 32);
                buf[i][j] = distij;
                buf[j][i] = distij;
            }
        int BestMedian = INT_MAX, BestIdx = 0;
        for (size_t i = 0; i < N; i++) {
            sort(buf[i].begin(), buf[i].end());
            const int median = buf[i][0.5 * (N - 1)];
            if (median < BestMedian) { BestMedian = median; BestIdx = (int)i; }
        }
        best_idx[m] = BestIdx; best_median[m] = BestMedian;
    }
}

// ---- KL_Track's optical flow (src/Frame.cpp:52-54): cv::calcOpticalFlowPyrLK with the reference's fixed arguments; see
// ora_lk.hpp for the restatement and how it is pinned. pts / next_pts are (x, y) pairs.
void ora_lk_track(const uint8_t* prev, const uint8_t* next, int w, int h, int stride, const float* pts, int n, float* next_pts, uint8_t* status,
                  float* err) {
    ora_lk::calc_lk(prev, next, w, h, stride, pts, n, next_pts, status, err);
}
// KL_Track (src/Frame.cpp:34-76) minus the MapPoint bookkeeping: kps = obj1->kps[GetMapPointIdx(mp)] in the order of
// GetMapPointsVector(). ok[i] = res[i] > 0 && err[i] < 1 (:57-58); new_kps[i] = what :65-69 would push onto obj2->kps (kps[i]
// with pt = next_pts[i], octave = 0), valid where ok[i]. The "fewer than 10 MapPoints -> return 0" gate (:41) applies: returns
// 0 and clears ok. Returns the number of ok points (the caller still skips MapPoints it has already seen, :61-63).
int ora_kl_track(const uint8_t* prev, const uint8_t* next, int w, int h, int stride, const KeyPoint* kps, int n, KeyPoint* new_kps, uint8_t* ok,
                 float* next_pts, uint8_t* status, float* err) {
    for (int i = 0; i < n; ++i) ok[i] = 0;
    if (n < 10) return 0;
    vector<float> pts((size_t)2 * n);
    for (int i = 0; i < n; ++i) { pts[2 * i] = kps[i].x; pts[2 * i + 1] = kps[i].y; }
    ora_lk::calc_lk(prev, next, w, h, stride, pts.data(), n, next_pts, status, err);
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        if (status[i] > 0 && err[i] < 1) {
            ok[i] = 1; ++cnt;
            new_kps[i] = kps[i]; new_kps[i].x = next_pts[2 * i]; new_kps[i].y = next_pts[2 * i + 1]; new_kps[i].octave = 0;
        }
    }
    return cnt;
}

// ---- Object::ComputeBow (src/Object.cpp:238-247) = DBoW3::Vocabulary::transform(features, BowVector&, FeatureVector&, levelsup)
// (modules/DBow3/src/Vocabulary.cpp:572-633) on a vocabulary passed as flat arrays: child_off/child_ids = m_nodes[i].children in
// stored order, node_desc = m_nodes[i].descriptor (32 B), word_id / weight = leaf fields, L = m_L. Per feature the tree descent
// of Vocabulary.cpp:641-672 (strict '<': the first of several equally close children wins), then BowVector::addWeight /
// addIfNotExist and FeatureVector::addFeature in feature order on real std::maps, then the /= size or BowVector::normalize.
// Outputs: per-feature word / weight / node, the BowVector as (ids ascending, values) and the FeatureVector flattened the way
// Matcher::DBowMatch consumes it. Returns the BowVector size; *n_fv = FeatureVector size.
int ora_bow_transform(const uint8_t* desc, int n, const int* child_off, const unsigned* child_ids, const uint8_t* node_desc, const int* word_id,
                      const double* weight, int L, int levelsup, int weighting, int norm_type /* 0 none, 1 L1, 2 L2 */, int* out_word, double* out_w,
                      unsigned* out_nid, unsigned* bow_ids, double* bow_vals, unsigned* fv_nodes, int* fv_off, int* fv_idx, int* n_fv) {
    map<unsigned, double> v;
    map<unsigned, vector<unsigned>> fv;
    const int nid_level = L - levelsup;
    for (int i = 0; i < n; ++i) {
        unsigned nid = 0, final_id = 0;   // nid_level <= 0: root
        int current_level = 0;
        do {
            ++current_level;
            double best_d = std::numeric_limits<double>::max();
            const unsigned node = final_id;
            for (int c = child_off[node]; c < child_off[node + 1]; ++c) {
                const unsigned id = child_ids[c];
                const double d = (double)hamming256(desc + (size_t)i * 32, node_desc + (size_t)id * 32);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (child_off[final_id] != child_off[final_id + 1]);
        const unsigned wid = (unsigned)word_id[final_id];
        const double w = weight[final_id];
        out_word[i] = (int)wid; out_w[i] = w; out_nid[i] = nid;
        if (w > 0) {
            if (weighting == 0 || weighting == 1) {   // TF_IDF, TF: addWeight
                auto it = v.lower_bound(wid);
                if (it != v.end() && !(wid < it->first)) it->second += w; else v.insert(it, make_pair(wid, w));
            } else {                                  // IDF, BINARY: addIfNotExist
                auto it = v.lower_bound(wid);
                if (it == v.end() || wid < it->first) v.insert(it, make_pair(wid, w));
            }
            fv[nid].push_back((unsigned)i);
        }
    }
    if ((weighting == 0 || weighting == 1) && !v.empty() && norm_type == 0) {
        const double nd = (double)v.size();
        for (auto& e : v) e.second /= nd;
    }
    if (norm_type != 0) {
        double norm = 0.0;
        if (norm_type == 1) { for (auto& e : v) norm += fabs(e.second); }
        else { for (auto& e : v) norm += e.second * e.second; norm = sqrt(norm); }
        if (norm > 0.0) for (auto& e : v) e.second /= norm;
    }
    int k = 0;
    for (auto& e : v) { bow_ids[k] = e.first; bow_vals[k] = e.second; ++k; }
    int m = 0, at = 0;
    for (auto& e : fv) { fv_nodes[m] = e.first; fv_off[m] = at; for (unsigned f : e.second) fv_idx[at++] = (int)f; ++m; }
    fv_off[m] = at;
    *n_fv = m;
    return k;
}

// ---- CPU baseline driver: n_frames three-camera frames (L, R, W images, each w*h contiguous), n_threads workers,
// each worker owns its three extractors (as Frame's statics) and runs extract x3 + stereo. Returns seconds. ----
double ora_bench_frames(const uint8_t* imgs, int n_frames, int w, int h, int nfeatures, float sf, int nlevels, int ini, int mn,
                        float bf, float b, int n_threads, int repeat, long long* total_kps) {
    atomic<int> next(0); atomic<long long> tk(0);
    const int total = n_frames * repeat;
    auto t0 = chrono::steady_clock::now();
    vector<thread> th;
    for (int t = 0; t < n_threads; ++t)
        th.emplace_back([&]() {
            Extractor E[3];
            for (auto& e : E) e.init(nfeatures, sf, nlevels, ini, mn);
            vector<KeyPoint> k[3]; vector<uint8_t> d[3]; vector<float> ur, dp;
            for (;;) {
                int f = next.fetch_add(1);
                if (f >= total) break;
                const uint8_t* base = imgs + (size_t)(f % n_frames) * 3 * w * h;
                for (int c = 0; c < 3; ++c) { k[c].clear(); Extract(E[c], Img{base + (size_t)c * w * h, w, h, (size_t)w}, k[c], d[c]); }
                ur.resize(k[0].size()); dp.resize(k[0].size());
                ComputeStereoMatch(E[0], E[1], k[0].data(), d[0].data(), (int)k[0].size(), k[1].data(), d[1].data(), (int)k[1].size(), h, bf, b, ur.data(), dp.data(), nullptr, nullptr);
                tk += (long long)(k[0].size() + k[1].size() + k[2].size());
            }
        });
    for (auto& x : th) x.join();
    double s = chrono::duration<double>(chrono::steady_clock::now() - t0).count();
    if (total_kps) *total_kps = tk.load();
    return s;
}
// CPU baseline for brute-force 2-NN over nq x nt pairs with n_threads (queries split), returns seconds.
double ora_bench_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int n_threads, DMatch* out) {
    auto t0 = chrono::steady_clock::now();
    vector<thread> th;
    for (int k = 0; k < n_threads; ++k)
        th.emplace_back([=]() {
            int a = (int)((long long)nq * k / n_threads), bnd = (int)((long long)nq * (k + 1) / n_threads);
            ora_knn2_bf(q + (size_t)a * 32, bnd - a, t, nt, out + (size_t)2 * a);
            for (int i = a; i < bnd; ++i) { out[2 * i].queryIdx = i; out[2 * i + 1].queryIdx = i; }
        });
    for (auto& x : th) x.join();
    return chrono::duration<double>(chrono::steady_clock::now() - t0).count();
}

}  // extern "C"
