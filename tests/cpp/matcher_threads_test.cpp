// Re-entrancy of the matcher side of the C ABI (include/mcv_b200.h, "Threading"): the reference's Matcher statics are stateless
// and callable from any thread (include/Matcher.hpp:58-92). N host threads hammer mcv_knn2_bf / mcv_knn2_firstparty /
// mcv_knn2_candidates / mcv_project_match / mcv_distinctive_descriptors concurrently, each on its own inputs and with sizes
// that force the scratch buffers to grow and shrink between calls; every result must equal the CPU oracle (the checker only).
//
//   matcher_threads_test [n_threads = 4] [rounds = 12]       exit 0 = all equal
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../../include/mcv_b200.h"

extern "C" {
int ora_knn2_bf(const uint8_t* q, int nq, const uint8_t* t, int nt, mcv_dmatch* out);
void ora_knn2_firstparty(const uint8_t* q, int nq, const uint8_t* t, int nt, mcv_dmatch* out);
void ora_knn2_candidates(const uint8_t* q, int nq, const uint8_t* t, const int* cand_off, const int* cand_idx, mcv_dmatch* out);
int ora_project_match(const mcv_keypoint* kps, const uint8_t* desc, int n, int W, int H, const float* scale_factors, const float* Rcw, const float* tcw,
                      const float* intr, const float* mp_xyz, const uint8_t* mp_desc, const int* mp_level, int n_mp, float r_threshold, int* out_idx,
                      int* out_dist);
void ora_distinctive(const uint8_t* desc, const int* off, int n_mp, int* best_idx, int* best_median);
}

struct Lcg {
    uint64_t s;
    explicit Lcg(uint64_t seed) : s(seed * 6364136223846793005ull + 1442695040888963407ull) {}
    uint32_t next() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 33); }
    int range(int lo, int hi) { return lo + (int)(next() % (uint32_t)(hi - lo)); }
    float uni(float lo, float hi) { return lo + (hi - lo) * (float)(next() & 0xffffff) / 16777216.f; }
};

static std::atomic<int> g_fail{0};
#define CHECK(cond, ...) do { if (!(cond)) { ++g_fail; fprintf(stderr, "FAIL thread %d round %d line %d: ", tid, round, __LINE__); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

static void desc_fill(std::vector<uint8_t>& d, int n, Lcg& r, bool low) {
    d.resize((size_t)n * 32);
    for (auto& b : d) b = (uint8_t)(low ? r.next() & 3 : r.next());
}

static void worker(int tid, int rounds) {
    Lcg rng(1000 + tid);
    for (int round = 0; round < rounds; ++round) {
        // sizes differ per thread and round: scratch buffers are re-grown while other threads are inside their kernels
        const int nq = rng.range(1, 1 + 400 * ((round + tid) % 5 + 1)), nt = rng.range(2, 2 + 900 * ((round * 3 + tid) % 4 + 1));
        std::vector<uint8_t> q, t;
        desc_fill(q, nq, rng, round & 1); desc_fill(t, nt, rng, round & 1);
        std::vector<mcv_dmatch> a((size_t)nq * 2), b((size_t)nq * 2);
        int k = 0;
        mcv_status st = mcv_knn2_bf(q.data(), nq, t.data(), nt, a.data(), &k);
        CHECK(st == MCV_OK, "knn2_bf status %d (%s)", st, mcv_last_error());
        ora_knn2_bf(q.data(), nq, t.data(), nt, b.data());
        CHECK(k == 2 && memcmp(a.data(), b.data(), a.size() * sizeof(mcv_dmatch)) == 0, "knn2_bf differs (nq %d nt %d)", nq, nt);
        st = mcv_knn2_firstparty(q.data(), nq, t.data(), nt, a.data());
        ora_knn2_firstparty(q.data(), nq, t.data(), nt, b.data());
        CHECK(st == MCV_OK && memcmp(a.data(), b.data(), a.size() * sizeof(mcv_dmatch)) == 0, "knn2_firstparty differs");
        std::vector<int> off(nq + 1, 0), cidx;
        for (int i = 0; i < nq; ++i) { const int len = rng.range(0, 12); for (int j = 0; j < len; ++j) cidx.push_back(rng.range(0, nt)); off[i + 1] = (int)cidx.size(); }
        st = mcv_knn2_candidates(q.data(), nq, t.data(), nt, off.data(), cidx.data(), a.data());
        ora_knn2_candidates(q.data(), nq, t.data(), off.data(), cidx.data(), b.data());
        CHECK(st == MCV_OK && memcmp(a.data(), b.data(), a.size() * sizeof(mcv_dmatch)) == 0, "knn2_candidates differs");
        // projection: nt keypoints spread over a 640 x 480 image, nq MapPoints near them
        std::vector<mcv_keypoint> kps(nt);
        for (int i = 0; i < nt; ++i) kps[i] = mcv_keypoint{rng.uni(20, 620), rng.uni(20, 460), 31.f, 0.f, 50.f, rng.range(0, 8), -1};
        float sf[8]; sf[0] = 1.f; for (int l = 1; l < 8; ++l) sf[l] = sf[l - 1] * 1.2f;
        const float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tc[3] = {0.01f, -0.02f, 0.03f}, K[4] = {500.f, 500.f, 320.f, 240.f};
        std::vector<float> xyz((size_t)nq * 3); std::vector<int> lvl(nq);
        for (int m = 0; m < nq; ++m) {
            const int src = rng.range(0, nt); const float z = rng.uni(2, 40);
            xyz[3 * m] = (kps[src].x - K[2]) / K[0] * z - tc[0]; xyz[3 * m + 1] = (kps[src].y - K[3]) / K[1] * z - tc[1]; xyz[3 * m + 2] = z - tc[2];
            lvl[m] = kps[src].octave;
            if (m % 3) memcpy(&q[(size_t)m * 32], &t[(size_t)src * 32], 32);
        }
        std::vector<int> oi(nq), od(nq), ri(nq), rd(nq);
        int cnt = 0;
        st = mcv_project_match(kps.data(), t.data(), nt, 640, 480, sf, 8, R, tc, K, xyz.data(), q.data(), lvl.data(), nq, 7.f, oi.data(), od.data(), &cnt);
        const int rc = ora_project_match(kps.data(), t.data(), nt, 640, 480, sf, R, tc, K, xyz.data(), q.data(), lvl.data(), nq, 7.f, ri.data(), rd.data());
        CHECK(st == MCV_OK && cnt == rc && oi == ri, "project_match differs (cnt %d vs %d)", cnt, rc);
        // distinctive descriptors: ragged observation lists cut out of t
        std::vector<int> moff(1, 0);
        while (moff.back() < nt) moff.push_back(std::min(nt, moff.back() + rng.range(0, 40)));
        const int n_mp = (int)moff.size() - 1;
        std::vector<int> bi(n_mp), bm(n_mp), ci(n_mp), cm(n_mp);
        st = mcv_distinctive_descriptors(t.data(), moff.data(), n_mp, bi.data(), bm.data(), nullptr);
        ora_distinctive(t.data(), moff.data(), n_mp, ci.data(), cm.data());
        CHECK(st == MCV_OK && bi == ci && bm == cm, "distinctive differs");
    }
}

int main(int argc, char** argv) {
    const int n_threads = argc > 1 ? atoi(argv[1]) : 4, rounds = argc > 2 ? atoi(argv[2]) : 12;
    if (mcv_device_count() < 1) { fprintf(stderr, "no CUDA device\n"); return 2; }
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(worker, i, rounds);
    for (auto& t : th) t.join();
    if (g_fail.load()) { fprintf(stderr, "%d check(s) failed\n", g_fail.load()); return 1; }
    printf("%d threads x %d rounds: all results equal the oracle\n", n_threads, rounds);
    return 0;
}
