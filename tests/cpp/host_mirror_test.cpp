// Drives the C++ host mirror (mcvslam_b200/host/mcvslam_b200.hpp) the way the reference's own test programs drive the
// reference (test/export_ORB_feature_extrac_result.cpp:47-59, test/matching_benchmark.cpp:37-62, test/test_stereo.cpp:28-46),
// and checks every result bit for bit against the CPU oracle (oracle/orb_oracle.cpp, linked here as the checker only).
//
//   host_mirror_test <tmp dir>             full check on cuda:0; exit 0 = all equal
//   host_mirror_test <tmp dir> --no-device asserts the no-CPU-fallback behaviour on a box without a GPU
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../mcvslam_b200/host/mcvslam_b200.hpp"

extern "C" {
void* ora_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void ora_orb_destroy(void* h);
int ora_orb_extract(void* h, const uint8_t* img, int w, int hgt, int stride, cv::KeyPoint* kps, int n_seeds, uint8_t* desc, int cap);
int ora_orb_level_size(void* h, int level, int* w, int* hgt);
void ora_orb_level_copy(void* h, int level, uint8_t* dst);
int ora_knn2_bf(const uint8_t* q, int nq, const uint8_t* t, int nt, cv::DMatch* out);
void ora_knn2_firstparty(const uint8_t* q, int nq, const uint8_t* t, int nt, cv::DMatch* out);
int ora_filter_ratio(const cv::DMatch* knn, int nq, int per, float ratio, cv::DMatch* out);
int ora_filter_threshold(cv::DMatch* m, int n, int thres_hold);
int ora_filter_orientation(cv::DMatch* m, int n, const cv::KeyPoint* kps1, const cv::KeyPoint* kps2);
int ora_stereo_match(void* hl, void* hr, const cv::KeyPoint* kl, const uint8_t* dl, int nl, const cv::KeyPoint* kr, const uint8_t* dr, int nr, int nRows,
                     float bf, float b, float* u_right, float* depth_left, int* best_dist, int* best_r);
int ora_project_match(const cv::KeyPoint* kps, const uint8_t* desc, int n, int W, int H, const float* scale_factors, const float* Rcw, const float* tcw,
                      const float* intr, const float* mp_xyz, const uint8_t* mp_desc, const int* mp_level, int n_mp, float r_threshold, int* out_idx,
                      int* out_dist);
int ora_fuse_match(const cv::KeyPoint* kps, const uint8_t* desc, int n, int W, int H, const float* sigma2, const float* inv_sigma2, const float* Rcw,
                   const float* tcw, const float* Ow, const float* intr, const float* depth_left, float bf, const float* mp_xyz, const float* mp_normal,
                   const uint8_t* mp_desc, const int* mp_level, int n_mp, int* out_idx, int* out_dist);
int ora_wnd_track(const cv::KeyPoint* kps1, const uint8_t* desc1, const int* q_idx, int n_q, const cv::KeyPoint* kps2, const uint8_t* desc2, int n2, int W,
                  int H, int* out_idx, int* out_best, int* out_dist);
void ora_distinctive(const uint8_t* desc, const int* off, int n_mp, int* best_idx, int* best_median);
int ora_kl_track(const uint8_t* prev, const uint8_t* next, int w, int h, int stride, const cv::KeyPoint* kps, int n, cv::KeyPoint* new_kps, uint8_t* ok,
                 float* next_pts, uint8_t* status, float* err);
int ora_distribute_octree(const cv::KeyPoint* in, int n, int minX, int maxX, int minY, int maxY, int N, cv::KeyPoint* out, int cap);
}

static int g_fail = 0;
#define CHECK(cond, ...)                                   \
    do {                                                   \
        if (!(cond)) {                                     \
            ++g_fail;                                      \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
            fprintf(stderr, __VA_ARGS__);                  \
            fprintf(stderr, "\n");                         \
        }                                                  \
    } while (0)

struct Lcg {
    uint64_t s;
    explicit Lcg(uint64_t seed) : s(seed * 6364136223846793005ull + 1442695040888963407ull) {}
    uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
    int range(int lo, int hi) { return lo + (int)(next() % (uint32_t)(hi - lo)); }
};

// random-rectangle scene + noise (same recipe as mcvslam_b200/synth.py, different generator: only determinism matters)
static cv::Mat scene(uint64_t seed, int w, int h) {
    Lcg r(seed);
    cv::Mat img(h, w, CV_8U);
    memset(img.data, 128, (size_t)w * h);
    for (int k = 0; k < w * h / 600; ++k) {
        const int x = r.range(0, w), y = r.range(0, h), rw = r.range(8, 64), rh = r.range(8, 64), g = r.range(0, 256);
        for (int yy = y; yy <= y + rh && yy < h; ++yy)
            for (int xx = x; xx <= x + rw && xx < w; ++xx) img.at<uint8_t>(yy, xx) = (uint8_t)g;
    }
    for (int i = 0; i < w * h; ++i) {
        const int v = img.data[i] + r.range(-6, 7);
        img.data[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
    return img;
}

static cv::Mat shifted_right(const cv::Mat& left, uint64_t seed) {
    const int w = left.cols, h = left.rows;
    cv::Mat fill = scene(seed + 77, w, h), right(h, w, CV_8U);
    for (int y = 0; y < h; ++y) {
        const int d = 4 + 56 * (y * 8 / h) / 7;
        for (int x = 0; x < w; ++x) right.at<uint8_t>(y, x) = x + d < w ? left.at<uint8_t>(y, x + d) : fill.at<uint8_t>(y, x);
    }
    return right;
}

static void write_file(const std::string& path, const std::string& text) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) { perror(path.c_str()); exit(2); }
    fputs(text.c_str(), f);
    fclose(f);
}

template <class T>
static bool same_bytes(const std::vector<T>& a, const std::vector<T>& b) { return a.size() == b.size() && (a.empty() || memcmp(a.data(), b.data(), a.size() * sizeof(T)) == 0); }

struct OracleOrb {
    void* h;
    std::vector<cv::KeyPoint> kps;
    std::vector<uint8_t> desc;
    OracleOrb(int nf, float sf, int nl, int ini, int mn) : h(ora_orb_create(nf, sf, nl, ini, mn)) {}
    ~OracleOrb() { ora_orb_destroy(h); }
    int extract(const cv::Mat& img, const std::vector<cv::KeyPoint>& seeds = {}) {
        const int cap = 4 * 4096;
        kps.assign(cap, cv::KeyPoint());
        for (size_t i = 0; i < seeds.size(); ++i) kps[i] = seeds[i];
        desc.assign((size_t)cap * 32, 0);
        const int n = ora_orb_extract(h, img.data, img.cols, img.rows, (int)img.step, kps.data(), (int)seeds.size(), desc.data(), cap);
        if (n >= 0) { kps.resize(n); desc.resize((size_t)n * 32); }
        return n;
    }
};

int main(int argc, char** argv) {
    using namespace MCVSLAM;
    const std::string tmp = argc > 1 ? argv[1] : "/tmp";
    const bool no_device = argc > 2 && std::string(argv[2]) == "--no-device";
    // the reference's config files, 8 levels as BASELINE.json configs[1] asks (shipped extractor.yaml has nlevels: 1)
    write_file(tmp + "/extractor.yaml",
               "nkeypoints: 2000\nscale_factor: 1.2\nnlevels: 8\n\n# ORB Extractor: Fast threshold\nORBextractor.iniThFAST: 28\nORBextractor.minThFAST: 15\n");
    write_file(tmp + "/frame.yaml",
               "Trl: 1. 0. 0. 1.0 0. 1. 0. 0. 0. 0. 1. 0. 0. 0. 0. 1.\nbf: 955.40503\nbaseline: 1.\n\nleft_extractor_path: \"${CURRENT_FOLDER}/extractor.yaml\"\n"
               "right_extractor_path: \"${CURRENT_FOLDER}/extractor.yaml\"\nwide_extractor_path: \"${CURRENT_FOLDER}/extractor.yaml\"\n");

    if (no_device) {
        ORB orb(tmp + "/extractor.yaml");
        CHECK(orb.status() == MCV_ERR_NO_DEVICE, "expected MCV_ERR_NO_DEVICE, got %d", orb.status());
        Keypoints kps; Desps desps;
        const int n = orb.Extract(scene(1, 640, 480), kps, desps);
        CHECK(n < -1, "Extract without a device must fail loudly, got %d", n);
        bool threw = false;
        try { Matcher::KnnMatch(cv::Mat(4, 32, CV_8U), cv::Mat(4, 32, CV_8U)); } catch (const std::runtime_error&) { threw = true; }
        CHECK(threw, "Matcher::KnnMatch without a device must throw");
        printf(g_fail ? "host_mirror_test: %d failure(s)\n" : "host_mirror_test: no-device behaviour ok\n", g_fail);
        return g_fail ? 1 : 0;
    }

    // ---- export_ORB_feature_extrac_result.cpp: one image -> Extract ----
    ORB orb(tmp + "/extractor.yaml");
    CHECK(orb.status() == MCV_OK, "ORB(config): %s", orb.last_error().c_str());
    CHECK(orb.GetLevels() == 8 && orb.GetScaleFactor() == 1.2f, "config parse");
    OracleOrb ref(2000, 1.2f, 8, 28, 15);
    const cv::Mat img = scene(11, 640, 480);
    Keypoints kps; Desps desps;
    const int n = orb.Extract(img, kps, desps);
    const int n_ref = ref.extract(img);
    CHECK(n == n_ref && n > 1900, "keypoint count %d vs oracle %d", n, n_ref);
    CHECK(same_bytes(kps, ref.kps), "keypoints differ from the oracle");
    CHECK(desps.rows == n && desps.cols == 32 && desps.isContinuous() && memcmp(desps.data, ref.desc.data(), (size_t)n * 32) == 0, "descriptors differ");
    CHECK((int)orb.mvImagePyramid.size() == 8, "mvImagePyramid levels");
    for (int l = 0; l < 8 && l < (int)orb.mvImagePyramid.size(); ++l) {
        int w = 0, h = 0;
        ora_orb_level_size(ref.h, l, &w, &h);
        std::vector<uint8_t> lv((size_t)w * h);
        ora_orb_level_copy(ref.h, l, lv.data());
        const cv::Mat& m = orb.mvImagePyramid[l];
        CHECK(m.cols == w && m.rows == h && memcmp(m.data, lv.data(), lv.size()) == 0, "mvImagePyramid[%d] differs", l);
    }
    const std::vector<float> sf = orb.GetScaleFactors(), is2 = orb.GetInverseScaleSigmaSquares();
    CHECK(sf.size() == 8 && sf[1] == 1.2f && is2[0] == 1.0f, "scale vectors");
    // empty image -> -1 (ORBextractor.cc:834); pre-seeded keypoints are kept (in/out kps)
    { Keypoints k; Desps d; CHECK(orb.Extract(cv::Mat(), k, d) == -1, "empty image must return -1"); }
    {
        Keypoints seeds = {cv::KeyPoint(100.4f, 80.6f, 31, 10.5f, 99, 0, 7), cv::KeyPoint(50.25f, 60.1f, 31, 359.9f, 99, 2, 7)};
        Keypoints k = seeds; Desps d;
        const int ns = orb.Extract(img, k, d);
        const int ns_ref = ref.extract(img, seeds);
        CHECK(ns == ns_ref && same_bytes(k, ref.kps) && memcmp(d.data, ref.desc.data(), (size_t)ns * 32) == 0, "pre-seeded extraction differs");
        ref.extract(img);
    }
    // static DistributeOctTree
    {
        Lcg r(5);
        Keypoints in;
        std::vector<uint8_t> used(608 * 448, 0);
        while (in.size() < 1500) {
            const int x = r.range(0, 608), y = r.range(0, 448);
            if (used[y * 608 + x]++) continue;
            in.push_back(cv::KeyPoint((float)x, (float)y, 7, -1, (float)r.range(15, 60), 0, -1));
        }
        Keypoints a = ORB::DistributeOctTree(in, 16, 16 + 608, 16, 16 + 448, 400, 0), b(in.size() + 16);
        b.resize(ora_distribute_octree(in.data(), (int)in.size(), 16, 16 + 608, 16, 16 + 448, 400, b.data(), (int)b.size()));
        CHECK(a.size() == b.size(), "DistributeOctTree count %zu vs %zu", a.size(), b.size());
        for (size_t i = 0; i < a.size() && i < b.size(); ++i)
            if (a[i].pt.x != b[i].pt.x || a[i].pt.y != b[i].pt.y || a[i].response != b[i].response) { CHECK(false, "DistributeOctTree differs at %zu", i); break; }
    }

    // ---- matching_benchmark.cpp: KnnMatch + filter chain between two frames ----
    ORB orb2(tmp + "/extractor.yaml");
    Keypoints kps2; Desps desps2;
    const cv::Mat img2 = scene(12, 640, 480);
    const int n2 = orb2.Extract(img2, kps2, desps2);
    {
        MatchResKnn knn = Matcher::KnnMatch(desps, desps2);
        std::vector<cv::DMatch> rk((size_t)n * 2);
        const int k = ora_knn2_bf(desps.data, n, desps2.data, n2, rk.data());
        CHECK(k == 2 && (int)knn.size() == n, "knn size");
        bool eq = true;
        for (int i = 0; i < n && eq; ++i) eq = knn[i].size() == 2 && memcmp(knn[i].data(), &rk[2 * (size_t)i], 32) == 0;
        CHECK(eq, "KnnMatch(Mat, Mat) differs from the oracle");
        MatchRes good = knn.FilterRatio().FilterThreshold();
        std::vector<cv::DMatch> rg((size_t)n);
        rg.resize(ora_filter_ratio(rk.data(), n, 2, 0.6f, rg.data()));
        rg.resize(ora_filter_threshold(rg.data(), (int)rg.size(), ORB_GOOD_THRESHOLD));
        CHECK(same_bytes(static_cast<std::vector<cv::DMatch>&>(good), rg), "FilterRatio().FilterThreshold() differs");
        MatchRes all = knn.FilterRatio(1.0f);
        std::vector<cv::DMatch> ra((size_t)n);
        ra.resize(ora_filter_ratio(rk.data(), n, 2, 1.0f, ra.data()));
        all.FilterOrientation(kps, kps2);
        ra.resize(ora_filter_orientation(ra.data(), (int)ra.size(), kps.data(), kps2.data()));
        CHECK(same_bytes(static_cast<std::vector<cv::DMatch>&>(all), ra), "FilterOrientation differs");
        MatchRes bf = Matcher::BFMatch(desps, desps2);
        bool eqb = (int)bf.size() == n;
        for (int i = 0; i < n && eqb; ++i) eqb = bf[i].trainIdx == rk[2 * (size_t)i].trainIdx && bf[i].distance == rk[2 * (size_t)i].distance;
        CHECK(eqb, "BFMatch differs");
        // the literal benchmark case: one query row vs (1 + N) train rows through the vector<Mat> overload
        std::vector<cv::Mat> q = {desps.row(1)}, t = {desps.row(1)};
        for (int i = 0; i < n2; ++i) t.push_back(desps2.row(i));
        MatchResKnn one = Matcher::KnnMatch(q, t);
        std::vector<uint8_t> tb(t.size() * 32);
        for (size_t i = 0; i < t.size(); ++i) memcpy(&tb[i * 32], t[i].data, 32);
        cv::DMatch r1[2];
        ora_knn2_firstparty(desps.row(1).data, 1, tb.data(), (int)t.size(), r1);
        CHECK(one.size() == 1 && one[0].size() == 2 && memcmp(one[0].data(), r1, 32) == 0 && one[0][0].trainIdx == 0 && one[0][0].distance == 0, "1 x (1+N) case");
        std::vector<cv::Mat> none;
        MatchResKnn pad = Matcher::KnnMatch(q, none);
        CHECK(pad[0][0].distance == 999 && pad[0][1].distance == 999 && pad[0][0].trainIdx == 0, "(0, 999) padding");
        CHECK(HammingDistance(desps.row(0), desps2.row(0)) == (uint)rk[0].distance || rk[0].trainIdx != 0, "HammingDistance");
    }

    // ---- test_stereo.cpp: one three-camera frame -> stereo depth ----
    Frame::Parse(tmp + "/frame.yaml");
    const cv::Mat left = scene(21, 640, 480), right = shifted_right(left, 21), wide = scene(23, 640, 480);
    Frame frame(left, right, wide);
    {
        OracleOrb ol(2000, 1.2f, 8, 28, 15), orr(2000, 1.2f, 8, 28, 15), ow(2000, 1.2f, 8, 28, 15);
        ol.extract(left); orr.extract(right); ow.extract(wide);
        CHECK(same_bytes(frame.LEFT->kps, ol.kps) && same_bytes(frame.RIGHT->kps, orr.kps) && same_bytes(frame.WIDE->kps, ow.kps), "Frame keypoints differ");
        const int nl = (int)ol.kps.size();
        std::vector<float> ur(nl), dp(nl);
        std::vector<int> bd(nl), br(nl);
        const int ns = ora_stereo_match(ol.h, orr.h, ol.kps.data(), ol.desc.data(), nl, orr.kps.data(), orr.desc.data(), (int)orr.kps.size(), 480, Frame::bf(),
                                        Frame::b(), ur.data(), dp.data(), bd.data(), br.data());
        CHECK(ns > 100, "synthetic stereo pair should match (%d)", ns);
        CHECK(same_bytes(frame.u_right, ur) && same_bytes(frame.depth_left, dp), "u_right / depth_left differ from the oracle");

        // ---- Tracker::Track's ProjectBunchMapPoints on the LEFT object ----
        Lcg r(9);
        std::vector<MapPointView> mps(3000);
        const float fx = 955.40503f * 640 / 512, cx = 320, cy = 240;
        std::vector<float> xyz; std::vector<uint8_t> md; std::vector<int> lvl;
        for (auto& mp : mps) {
            const int src = r.range(0, nl);
            const float z = 2.f + (float)r.range(0, 4800) / 100.f;
            const float u = ol.kps[src].pt.x + (float)r.range(-30, 31) / 10.f, v = ol.kps[src].pt.y + (float)r.range(-30, 31) / 10.f;
            mp.xyz[0] = (u - cx) / fx * z - 0.02f; mp.xyz[1] = (v - cy) / fx * z + 0.01f; mp.xyz[2] = z - 0.03f;
            memcpy(mp.desp, &ol.desc[(size_t)src * 32], 32);
            for (int f = r.range(0, 30); f > 0; --f) { const int b = r.range(0, 256); mp.desp[b >> 3] ^= (uint8_t)(1 << (b & 7)); }
            mp.level = ol.kps[src].octave;
            xyz.insert(xyz.end(), mp.xyz, mp.xyz + 3); md.insert(md.end(), mp.desp, mp.desp + 32); lvl.push_back(mp.level);
        }
        const float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0.02f, -0.01f, 0.03f};
        frame.LEFT->SetPose(R, t);
        frame.LEFT->SetIntrinsics(fx, fx, cx, cy);
        for (float r_th : {5.f, 10.f}) {
            std::vector<int> mi, mdist, oi(mps.size()), od(mps.size());
            const uint cnt = frame.LEFT->ProjectBunchMapPoints(mps, r_th, mi, &mdist);
            const int cnt_ref = ora_project_match(ol.kps.data(), ol.desc.data(), nl, 640, 480, frame.LEFT->extractor->mvScaleFactor.data(), R, t,
                                                  frame.LEFT->intr, xyz.data(), md.data(), lvl.data(), (int)mps.size(), r_th, oi.data(), od.data());
            CHECK((int)cnt == cnt_ref && cnt > 500 && mi == oi && mdist == od, "ProjectBunchMapPoints(r=%g): %u vs %d", r_th, cnt, cnt_ref);
        }

        // ---- Map::Fuse front-end on the LEFT object (kf = this frame: its depth_left / bf), Tracker::Wnd_Track LEFT -> WIDE ----
        {
            std::vector<float> nrm;
            const float Ow[3] = {-t[0], -t[1], -t[2]};   // R = I
            for (auto& mp : mps) {
                float v[3] = {mp.xyz[0] - Ow[0], mp.xyz[1] - Ow[1], mp.xyz[2] - Ow[2]};
                const float len = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
                for (int a = 0; a < 3; ++a) mp.normal[a] = v[a] / len;
                if (r.range(0, 10) == 0) mp.normal[2] = -mp.normal[2];   // some back-facing
                nrm.insert(nrm.end(), mp.normal, mp.normal + 3);
            }
            std::vector<int> mi, mdist, oi(mps.size()), od(mps.size());
            const uint cnt = frame.LEFT->FuseMatch(mps, frame.depth_left, Frame::bf(), mi, &mdist);
            const int cnt_ref = ora_fuse_match(ol.kps.data(), ol.desc.data(), nl, 640, 480, frame.LEFT->extractor->mvLevelSigma2.data(),
                                               frame.LEFT->extractor->mvInvLevelSigma2.data(), R, t, Ow, frame.LEFT->intr, frame.depth_left.data(), Frame::bf(),
                                               xyz.data(), nrm.data(), md.data(), lvl.data(), (int)mps.size(), oi.data(), od.data());
            CHECK((int)cnt == cnt_ref && cnt > 100 && mi == oi && mdist == od, "FuseMatch: %u vs %d", cnt, cnt_ref);
            std::vector<int> q;
            for (int i = 0; i < nl; i += 3) q.push_back(i);
            std::vector<int> wm, wd, wi(q.size()), wb(q.size()), wdd(q.size());
            const uint wcnt = Wnd_Track(*frame.LEFT, q, *frame.RIGHT, wm, &wd);
            const int wref = ora_wnd_track(ol.kps.data(), ol.desc.data(), q.data(), (int)q.size(), orr.kps.data(), orr.desc.data(), (int)orr.kps.size(), 640, 480,
                                           wi.data(), wb.data(), wdd.data());
            CHECK((int)wcnt == wref && wm == wi && wd == wdd, "Wnd_Track: %u vs %d", wcnt, wref);
        }

        // ---- MapPoint::ComputeDistinctiveDescriptors on a MapPoint observed by every 40th LEFT keypoint; KL_Track LEFT -> RIGHT ----
        {
            std::vector<cv::Mat> obs;
            std::vector<uint8_t> rows;
            for (int i = 0; i < nl; i += 40) { obs.push_back(frame.LEFT->desps.row(i)); rows.insert(rows.end(), ol.desc.begin() + (size_t)i * 32, ol.desc.begin() + (size_t)i * 32 + 32); }
            int bi = -1, bm = -1, rbi = -1, rbm = -1;
            const cv::Mat best = ComputeDistinctiveDescriptors(obs, &bi, &bm);
            const int off[2] = {0, (int)obs.size()};
            ora_distinctive(rows.data(), off, 1, &rbi, &rbm);
            CHECK(bi == rbi && bm == rbm && memcmp(best.data, rows.data() + (size_t)bi * 32, 32) == 0, "ComputeDistinctiveDescriptors: %d/%d vs %d/%d", bi, bm, rbi, rbm);
            std::vector<int> q;
            for (int i = 0; i < nl; i += 4) q.push_back(i);
            std::vector<cv::KeyPoint> sel(q.size()), rk(q.size());
            for (size_t i = 0; i < q.size(); ++i) sel[i] = ol.kps[(size_t)q[i]];
            std::vector<uint8_t> seen(q.size(), 0), rok(q.size()), rst(q.size());
            seen[3] = 1;                                                  // a MapPoint an earlier frame already tracked (src/Frame.cpp:61-63)
            std::vector<float> rnx(2 * q.size()), rerr(q.size());
            ora_kl_track(left.data, right.data, 640, 480, 640, sel.data(), (int)q.size(), rk.data(), rok.data(), rnx.data(), rst.data(), rerr.data());
            Object target(right, nullptr);
            target.kps = orr.kps;
            std::vector<int> new_idx;
            const uint kcnt = KL_Track(*frame.LEFT, q, target, seen, new_idx);
            uint want = 0; bool same = true; size_t at = orr.kps.size();
            for (size_t i = 0; i < q.size(); ++i) {
                if (!rok[i] || i == 3) { same &= new_idx[i] == -1; continue; }
                same &= new_idx[i] == (int)at && memcmp(&target.kps[at], &rk[i], 28) == 0;
                ++at; ++want;
            }
            CHECK(kcnt == want && same && target.kps.size() == at, "KL_Track: %u vs %u", kcnt, want);
        }

        // ---- batched rig == per-frame Frame ----
        Rig rig(2000, 1.2f, 8, 28, 15, Frame::bf(), Frame::b());
        std::vector<uint8_t> imgs((size_t)2 * 3 * 640 * 480);
        const cv::Mat* trip[3] = {&left, &right, &wide};
        for (int f = 0; f < 2; ++f)
            for (int c = 0; c < 3; ++c) memcpy(&imgs[((size_t)f * 3 + c) * 640 * 480], trip[c]->data, (size_t)640 * 480);
        std::vector<cv::KeyPoint> bk; std::vector<uint8_t> bdsc; std::vector<int32_t> cnts; std::vector<float> bur, bdp;
        rig.Process(imgs.data(), 2, 640, 480, bk, bdsc, cnts, bur, bdp);
        for (int f = 0; f < 2; ++f) {
            CHECK(cnts[3 * f] == nl, "rig count");
            CHECK(memcmp(&bk[(size_t)(3 * f) * rig.cap()], ol.kps.data(), (size_t)nl * 28) == 0, "rig keypoints (frame %d)", f);
            CHECK(memcmp(&bur[(size_t)f * rig.cap()], ur.data(), (size_t)nl * 4) == 0 && memcmp(&bdp[(size_t)f * rig.cap()], dp.data(), (size_t)nl * 4) == 0, "rig stereo (frame %d)", f);
        }
    }
    printf(g_fail ? "host_mirror_test: %d failure(s)\n" : "host_mirror_test: all results equal the oracle\n", g_fail);
    return g_fail ? 1 : 0;
}
