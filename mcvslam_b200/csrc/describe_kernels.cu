// Orientation, steered rBRIEF descriptor and final keypoint assembly for a batch of images (sm_100a).
//
// One warp per output keypoint. Replaces, for every keypoint:
//   * the post-distribution fix-up (ORBextractor.cc:640-649): pt += 16, octave, size = (int)(31 * scale);
//   * IC_Angle (ORBextractor.cc:75-98): int32 moments over the 15-px circular patch -> cv::fastAtan2;
//   * computeOrbDescriptor (ORBextractor.cc:101-141) on the 7x7-Gaussian-smoothed level: 256 steered point pairs,
//     lane i produces descriptor byte i;
//   * the level-major assembly of operator() (ORBextractor.cc:845-897): quadtree keypoints of a level in heap-pop order,
//     then the caller's pre-seeded keypoints of that octave, pt *= scale for level != 0.
#include "devmath.cuh"
#include "engine.h"

namespace mcv {

__constant__ int8_t c_pattern[1024] = {
#include "rbrief_pattern.inc"
};

constexpr int DESC_WARPS = 8;

__global__ void __launch_bounds__(32 * DESC_WARPS) k_orient_desc(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                                                                 const uint32_t* __restrict__ out_pts, const int* __restrict__ out_cnt,
                                                                 const mcv_keypoint* __restrict__ seeds, int n_seeds,
                                                                 mcv_keypoint* __restrict__ kps, uint8_t* __restrict__ desc,
                                                                 int* __restrict__ counts, int cap, const __grid_constant__ Plan P) {
    // pattern transposed into shared memory: s_pat[k][lane] = point (16*lane + k) as (x, y) -> conflict-free per-lane reads
    __shared__ char2 s_pat[16][32];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_pat[i & 15][i >> 4] = make_char2(c_pattern[2 * i], c_pattern[2 * i + 1]);
    __syncthreads();
    const int img = blockIdx.y, lane = threadIdx.x & 31;
    const int slot = blockIdx.x * DESC_WARPS + (threadIdx.x >> 5);
    // locate (level, index within level) of this output slot
    int level = -1, j = 0, base = 0, n_quad = 0;
    for (int l = 0; l < P.n_levels; ++l) {
        const int nq = out_cnt[(size_t)img * P.n_levels + l];
        int ns = 0;
        for (int s = 0; s < n_seeds; ++s) ns += seeds[s].octave == l;  // n_seeds is 0 on the batch path
        if (level < 0 && slot < base + nq + ns) { level = l; j = slot - base; n_quad = nq; }
        base += nq + ns;
    }
    if (slot == 0 && lane == 0) counts[img] = min(base, cap);
    if (level < 0 || slot >= cap) return;
    const LevelGeom& g = P.lv[level];
    const uint8_t* im = pyr + (size_t)img * P.pyr_bytes + g.img_off;
    const uint8_t* bl = blur + (size_t)img * P.pyr_bytes + g.img_off;

    mcv_keypoint kp;
    int cx, cy;
    if (j < n_quad) {
        const uint32_t p = out_pts[(size_t)img * P.out_per_image + g.out_off + j];
        cx = pt_x(p) + BORDER; cy = pt_y(p) + BORDER;
        kp.x = (float)cx; kp.y = (float)cy;
        kp.size = (float)g.kp_size; kp.response = (float)pt_r(p); kp.octave = level; kp.class_id = -1;
        // IC_Angle: lane <-> column u = lane - 15; rows v = -15..15; |u| <= umax[|v|]
        int m10 = 0, m01 = 0;
        const int u = lane - 15;
        const uint8_t* c = im + (size_t)cy * g.pitch + cx;
        constexpr int UMAX[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
#pragma unroll
        for (int v = -15; v <= 15; ++v) {
            const int d = UMAX[v < 0 ? -v : v];
            if (lane < 31 && u >= -d && u <= d) {
                const int val = c[v * g.pitch + u];
                m10 += u * val;
                m01 += v * val;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
        kp.angle = fast_atan2_deg((float)m01, (float)m10);
    } else {
        // pre-seeded keypoint: the (j - n_quad)-th seed of this octave, in the caller's order (ORBextractor.cc:845-847)
        int k = j - n_quad, s = 0;
        for (; s < n_seeds; ++s) if (seeds[s].octave == level && k-- == 0) break;
        kp = seeds[s];
        cx = cv_round_f(kp.x); cy = cv_round_f(kp.y);
    }
    // steered BRIEF
    const float ang = __fmul_rn(kp.angle, 0.017453292519943295f);  // factorPI = (float)(CV_PI / 180.f)
    float a, b;
    sincosf_glibc(ang, &b, &a);  // a = cos, b = sin
    const uint8_t* center = bl + (size_t)cy * g.pitch + cx;
    unsigned val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const char2 p0 = s_pat[2 * k][lane], p1 = s_pat[2 * k + 1][lane];
        const float x0 = (float)p0.x, y0 = (float)p0.y, x1 = (float)p1.x, y1 = (float)p1.y;
        const int r0 = cv_round_f(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a))), c0 = cv_round_f(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = cv_round_f(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a))), c1 = cv_round_f(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = center[r0 * g.pitch + c0], t1 = center[r1 * g.pitch + c1];
        val |= (unsigned)(t0 < t1) << k;
    }
    desc[((size_t)img * cap + slot) * 32 + lane] = (uint8_t)val;
    if (level != 0) { kp.x = __fmul_rn(kp.x, g.scale); kp.y = __fmul_rn(kp.y, g.scale); }
    if (lane == 0) kps[(size_t)img * cap + slot] = kp;
}

int launch_orient_desc(const Plan& P, const uint8_t* d_pyr, const uint8_t* d_blur, const uint32_t* d_out_pts, const int* d_out_cnt,
                       const SeedInfo* seeds, mcv_keypoint* d_kps, uint8_t* d_desc, int* d_counts, int cap, int n_images,
                       cudaStream_t s) {
    const int n_seeds = seeds ? seeds->n_seeds : 0;
    const int max_kp = std::min(cap, P.max_quad_kp + n_seeds);
    dim3 grid((max_kp + DESC_WARPS - 1) / DESC_WARPS, n_images);
    k_orient_desc<<<grid, 32 * DESC_WARPS, 0, s>>>(d_pyr, d_blur, d_out_pts, d_out_cnt, seeds ? seeds->d_seeds : nullptr, n_seeds, d_kps,
                                                  d_desc, d_counts, cap, P);
    return 1;
}

}  // namespace mcv
