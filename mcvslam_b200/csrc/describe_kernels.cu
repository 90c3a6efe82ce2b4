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
constexpr int DESC_KPW = 4;      // keypoints per warp (amortises the per-CTA tables and the level lookup)

__device__ __forceinline__ int dp4a_us(unsigned a, unsigned b_signed, int c) {   // sum of u8(a) * s8(b) + c
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_signed), "r"(c));
    return d;
}

__global__ void __launch_bounds__(32 * DESC_WARPS) k_orient_desc(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                                                                 const uint32_t* __restrict__ out_pts, const int* __restrict__ out_cnt,
                                                                 const mcv_keypoint* __restrict__ seeds, int n_seeds,
                                                                 mcv_keypoint* __restrict__ kps, uint8_t* __restrict__ desc,
                                                                 int* __restrict__ counts, int cap, const __grid_constant__ Plan P) {
    // pattern transposed into shared memory as floats: s_pat[k][lane] = point (16*lane + k) -> conflict-free per-lane reads
    __shared__ float2 s_pat[16][32];
    // IC_Angle weights of the 31x31 circular patch, by row (v + 15) and 4-px word k (u = -15 + 4k ...): byte = u inside the
    // circle (|u| <= umax[|v|], ORBextractor.cc:444-456), 0 outside; s_one has 1 / 0. Row 31 is all zero.
    __shared__ unsigned s_wu[32][8], s_one[32][8];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_pat[i & 15][i >> 4] = make_float2((float)c_pattern[2 * i], (float)c_pattern[2 * i + 1]);
    {
        constexpr int UMAX[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
        const int row = threadIdx.x >> 3, k = threadIdx.x & 7;          // 256 threads <-> 32 x 8 entries
        unsigned wu = 0, one = 0;
        if (row < 31) {
            const int v = row - 15, d = UMAX[v < 0 ? -v : v];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int u = -15 + 4 * k + j;
                if (u >= -d && u <= d) { wu |= (unsigned)(u & 0xff) << (8 * j); one |= 1u << (8 * j); }
            }
        }
        s_wu[row][k] = wu; s_one[row][k] = one;
    }
    __syncthreads();
    const int img = blockIdx.y, lane = threadIdx.x & 31;
    // per-level slot ranges of this image: quadtree keypoints of the level, then the caller's seeds of that octave
    int my_nq = 0, my_ns = 0;
    if (lane < P.n_levels) {
        my_nq = out_cnt[(size_t)img * P.n_levels + lane];
        for (int s = 0; s < n_seeds; ++s) my_ns += seeds[s].octave == lane;   // n_seeds is 0 on the batch path
    }
    int my_end = my_nq + my_ns;                                         // inclusive prefix = first slot after this level
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, my_end, o); if (lane >= o) my_end += v; }
    const int total = __shfl_sync(0xffffffffu, my_end, 31);
    const int warp_slot0 = (blockIdx.x * DESC_WARPS + (threadIdx.x >> 5)) * DESC_KPW;
    if (warp_slot0 == 0 && lane == 0) counts[img] = min(total, cap);

    for (int q = 0; q < DESC_KPW; ++q) {
        const int slot = warp_slot0 + q;
        if (slot >= total || slot >= cap) return;                       // warp-uniform
        const int level = __popc(__ballot_sync(0xffffffffu, lane < P.n_levels && slot >= my_end));
        const int lvl_end = __shfl_sync(0xffffffffu, my_end, level);
        const int n_quad = __shfl_sync(0xffffffffu, my_nq, level), n_sd = __shfl_sync(0xffffffffu, my_ns, level);
        const int j = slot - (lvl_end - n_quad - n_sd);
        const LevelGeom& g = P.lv[level];
        const int pitch = g.pitch;
        const uint8_t* im = pyr + (size_t)img * P.pyr_bytes + g.img_off;
        const uint8_t* bl = blur + (size_t)img * P.pyr_bytes + g.img_off;

        mcv_keypoint kp;
        int cx, cy;
        if (j < n_quad) {
            const uint32_t p = out_pts[(size_t)img * P.out_per_image + g.out_off + j];
            cx = pt_x(p) + BORDER; cy = pt_y(p) + BORDER;
            kp.x = (float)cx; kp.y = (float)cy;
            kp.size = (float)g.kp_size; kp.response = (float)pt_r(p); kp.octave = level; kp.class_id = -1;
            // IC_Angle: lane = (row group r4, word k); 8 rounds of 4 rows. The row's 31 patch bytes start at alignment `a` of the
            // first aligned word (level rows are 4-byte aligned), so lane k funnel-shifts words k, k+1 into its 4 patch pixels.
            const int r4 = lane >> 3, k = lane & 7;
            const int a8 = ((cx - 15) & 3) * 8;
            const uint8_t* p0 = im + (size_t)(cy - 15) * pitch + ((cx - 15) & ~3) + 4 * k;
            int m10 = 0, m01 = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = 4 * i + r4;
                const uint8_t* pr = p0 + min(row, 30) * pitch;          // row 31 has zero weights
                const unsigned w0 = __ldg(reinterpret_cast<const unsigned*>(pr)), w1 = __ldg(reinterpret_cast<const unsigned*>(pr + 4));
                const unsigned w = __funnelshift_r(w0, w1, a8);
                m10 = dp4a_us(w, s_wu[row][k], m10);
                m01 += (row - 15) * (int)__dp4a(w, s_one[row][k], 0u);
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
        const uint8_t* center = bl + (size_t)cy * pitch + cx;
        unsigned val = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float2 q0 = s_pat[2 * k][lane], q1 = s_pat[2 * k + 1][lane];
            const int r0 = cv_round_f(__fadd_rn(__fmul_rn(q0.x, b), __fmul_rn(q0.y, a))), c0 = cv_round_f(__fsub_rn(__fmul_rn(q0.x, a), __fmul_rn(q0.y, b)));
            const int r1 = cv_round_f(__fadd_rn(__fmul_rn(q1.x, b), __fmul_rn(q1.y, a))), c1 = cv_round_f(__fsub_rn(__fmul_rn(q1.x, a), __fmul_rn(q1.y, b)));
            const int t0 = center[r0 * pitch + c0], t1 = center[r1 * pitch + c1];
            val |= (unsigned)(t0 < t1) << k;
        }
        desc[((size_t)img * cap + slot) * 32 + lane] = (uint8_t)val;
        if (level != 0) { kp.x = __fmul_rn(kp.x, g.scale); kp.y = __fmul_rn(kp.y, g.scale); }
        if (lane == 0) kps[(size_t)img * cap + slot] = kp;
    }
}

int launch_orient_desc(const Plan& P, const uint8_t* d_pyr, const uint8_t* d_blur, const uint32_t* d_out_pts, const int* d_out_cnt,
                       const SeedInfo* seeds, mcv_keypoint* d_kps, uint8_t* d_desc, int* d_counts, int cap, int n_images,
                       cudaStream_t s) {
    const int n_seeds = seeds ? seeds->n_seeds : 0;
    const int max_kp = std::max(1, std::min(cap, P.max_quad_kp + n_seeds));
    constexpr int per_cta = DESC_WARPS * DESC_KPW;
    dim3 grid((max_kp + per_cta - 1) / per_cta, n_images);
    k_orient_desc<<<grid, 32 * DESC_WARPS, 0, s>>>(d_pyr, d_blur, d_out_pts, d_out_cnt, seeds ? seeds->d_seeds : nullptr, n_seeds, d_kps,
                                                  d_desc, d_counts, cap, P);
    return 1;
}

}  // namespace mcv
