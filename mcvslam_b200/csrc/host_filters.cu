// Host-side, order-defining epilogues of the matcher (src/Matcher.cpp:23-35,44-74,76-91,100-111) and the BoW-guided
// matcher front-end (src/Matcher.cpp:146-193). These are O(matches) sequential passes whose OUTPUT ORDER is part of the
// reference's behaviour (swap-remove, std::sort of histogram bins), so they stay on the host and operate on the 2-NN
// results the GPU kernels produce. The Hamming work of DBowMatch runs in k_knn2_candidates.
#include <math.h>

#include <algorithm>
#include <vector>

#include "engine.h"

extern "C" {

mcv_status mcv_filter_ratio(const mcv_dmatch* knn, int nq, int per, float ratio, mcv_dmatch* out, int* n_out) {
    if (nq < 0 || per < 1 || !n_out || (nq > 0 && (!knn || !out))) return MCV_ERR_BAD_ARG;
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        const mcv_dmatch* row = knn + (size_t)per * i;
        if (per == 1) out[n++] = row[0];
        else if (row[0].distance / row[1].distance <= ratio) out[n++] = row[0];
    }
    *n_out = n;
    return MCV_OK;
}

mcv_status mcv_filter_threshold(mcv_dmatch* m, int* n_io, int thres_hold) {
    if (!n_io || *n_io < 0 || (*n_io > 0 && !m)) return MCV_ERR_BAD_ARG;
    int i = 0, j = *n_io - 1;
    while (i <= j) {
        if (m[i].distance > thres_hold) m[i] = m[j--];
        else i++;
    }
    *n_io = i;
    return MCV_OK;
}

mcv_status mcv_filter_orientation(mcv_dmatch* m, int* n_io, const mcv_keypoint* kps1, int n1, const mcv_keypoint* kps2, int n2) {
    if (!n_io || *n_io < 0 || (*n_io > 0 && (!m || !kps1 || !kps2))) return MCV_ERR_BAD_ARG;
    const int HISTO_LENGTH = 40;
    const float HISTO_FACTOR = 1.0f / (360.0f / HISTO_LENGTH);
    std::vector<unsigned> rot_bins[HISTO_LENGTH];
    for (int i = 0; i < HISTO_LENGTH; i++) rot_bins[i].reserve(500);
    for (unsigned i = 0, sz = (unsigned)*n_io; i < sz; i++) {
        if (m[i].queryIdx < 0 || m[i].queryIdx >= n1 || m[i].trainIdx < 0 || m[i].trainIdx >= n2) return MCV_ERR_BAD_ARG;
        float rot = kps1[m[i].queryIdx].angle - kps2[m[i].trainIdx].angle;
        if (rot < 0.0) rot += 360.0f;
        int bin_id = (int)roundf(rot * HISTO_FACTOR);
        if (bin_id == HISTO_LENGTH) bin_id = 0;
        if (bin_id < 0 || bin_id >= HISTO_LENGTH) return MCV_ERR_BAD_ARG;  // the reference asserts here
        rot_bins[bin_id].push_back(i);
    }
    std::sort(&rot_bins[0], &rot_bins[0] + HISTO_LENGTH,
              [](const std::vector<unsigned>& a, const std::vector<unsigned>& b) { return a.size() > b.size(); });
    std::vector<mcv_dmatch> ret;
    for (unsigned i = 0; i < 3; i++)
        for (unsigned idx : rot_bins[i]) ret.push_back(m[idx]);
    for (size_t i = 0; i < ret.size(); ++i) m[i] = ret[i];
    *n_io = (int)ret.size();
    return MCV_OK;
}

mcv_status mcv_filter_fmatrix(mcv_dmatch* m, int* n_io, const mcv_keypoint* kps1, int n1, const mcv_keypoint* kps2, int n2, const float* F,
                              const float* level_sigma2, int nlevels) {
    if (!n_io || *n_io < 0 || !F || !level_sigma2 || (*n_io > 0 && (!m || !kps1 || !kps2))) return MCV_ERR_BAD_ARG;
    int i = 0, j = *n_io - 1;
    while (i <= j) {
        const mcv_dmatch d = m[i];
        if (d.queryIdx < 0 || d.queryIdx >= n1 || d.trainIdx < 0 || d.trainIdx >= n2) return MCV_ERR_BAD_ARG;
        const mcv_keypoint& kp1 = kps1[d.queryIdx];
        const mcv_keypoint& kp2 = kps2[d.trainIdx];
        if (kp2.octave < 0 || kp2.octave >= nlevels) return MCV_ERR_BAD_ARG;
        // CheckDistEpipolarLine: l = x1' F12 = [a b c]
        const float a = kp1.x * F[0] + kp1.y * F[3] + F[6];
        const float b = kp1.x * F[1] + kp1.y * F[4] + F[7];
        const float c = kp1.x * F[2] + kp1.y * F[5] + F[8];
        const float num = a * kp2.x + b * kp2.y + c;
        const float den = a * a + b * b;
        bool ok = false;
        if (den != 0) {
            const float dsqr = num * num / den;
            ok = dsqr < 3.84 * level_sigma2[kp2.octave];
        }
        if (!ok) m[i] = m[j--];
        else i++;
    }
    *n_io = i;
    return MCV_OK;
}

mcv_status mcv_dbow_match(const uint8_t* desc1, int n1, const uint32_t* node_ids1, const int32_t* feat_off1, const int32_t* feat_idx1, int n_nodes1,
                          const uint8_t* desc2, int n2, const uint32_t* node_ids2, const int32_t* feat_off2, const int32_t* feat_idx2, int n_nodes2,
                          mcv_dmatch* out, int* n_pairs) {
    if (!n_pairs || n1 < 0 || n2 < 0 || n_nodes1 < 0 || n_nodes2 < 0) return MCV_ERR_BAD_ARG;
    *n_pairs = 0;
    if (n_nodes1 == 0 || n_nodes2 == 0) return MCV_OK;
    if (!desc1 || !desc2 || !node_ids1 || !node_ids2 || !feat_off1 || !feat_off2 || !feat_idx1 || !feat_idx2 || !out) return MCV_ERR_BAD_ARG;
    // merge-join over the two ordered maps; each query (feature of image 1 in a shared node) gets the node's image-2 features
    // as its candidate list, in the map's order
    std::vector<int32_t> q_feat, off(1, 0), cidx;
    int a = 0, b = 0;
    while (a < n_nodes1 && b < n_nodes2) {
        if (node_ids1[a] == node_ids2[b]) {
            for (int i = feat_off1[a]; i < feat_off1[a + 1]; ++i) {
                if (feat_idx1[i] < 0 || feat_idx1[i] >= n1) return MCV_ERR_BAD_ARG;
                q_feat.push_back(feat_idx1[i]);
                for (int j = feat_off2[b]; j < feat_off2[b + 1]; ++j) {
                    if (feat_idx2[j] < 0 || feat_idx2[j] >= n2) return MCV_ERR_BAD_ARG;
                    cidx.push_back(feat_idx2[j]);
                }
                off.push_back((int32_t)cidx.size());
            }
            ++a; ++b;
        } else if (node_ids1[a] < node_ids2[b]) {
            a = (int)(std::lower_bound(node_ids1 + a, node_ids1 + n_nodes1, node_ids2[b]) - node_ids1);
        } else {
            b = (int)(std::lower_bound(node_ids2 + b, node_ids2 + n_nodes2, node_ids1[a]) - node_ids2);
        }
    }
    const int nq = (int)q_feat.size();
    if (nq == 0) return MCV_OK;
    std::vector<uint8_t> q((size_t)nq * 32);
    for (int i = 0; i < nq; ++i) memcpy(&q[(size_t)i * 32], desc1 + (size_t)q_feat[i] * 32, 32);
    std::vector<mcv_dmatch> knn((size_t)nq * 2);
    mcv_status st = mcv_knn2_candidates(q.data(), nq, desc2, n2, off.data(), cidx.data(), knn.data());
    if (st) return st;
    int n = 0;
    for (int i = 0; i < nq; ++i) {
        if (knn[2 * i + 1].distance != 999.f) {  // `if (d[1] != 999)` — needs a second neighbour
            for (int e = 0; e < 2; ++e) {
                const mcv_dmatch& k = knn[2 * i + e];
                out[2 * n + e] = mcv_dmatch{q_feat[i], cidx[off[i] + k.trainIdx], -1, k.distance};
            }
            ++n;
        }
    }
    *n_pairs = n;
    return MCV_OK;
}

}  // extern "C"
