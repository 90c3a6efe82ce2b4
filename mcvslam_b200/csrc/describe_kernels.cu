// Orientation, steered rBRIEF descriptor and final keypoint assembly for a batch of images (sm_100a).
//
// A quarter-warp per output keypoint for the scalar work and the orientation, the whole warp per keypoint for the descriptor
// (see "Design" below). Replaces, for every keypoint:
//   * the post-distribution fix-up (ORBextractor.cc:640-649): pt += 16, octave, size = (int)(31 * scale);
//   * IC_Angle (ORBextractor.cc:75-98): int32 moments over the 15-px circular patch -> cv::fastAtan2;
//   * computeOrbDescriptor (ORBextractor.cc:101-141) on the 7x7-Gaussian-smoothed level: 256 steered point pairs,
//     lane b of the warp produces descriptor byte b of each of the warp's four keypoints in turn;
//   * the level-major assembly of operator() (ORBextractor.cc:845-897): quadtree keypoints of a level in heap-pop order,
//     then the caller's pre-seeded keypoints of that octave, pt *= scale for level != 0.
#include <string.h>
#include "devmath.cuh"
#include "tma.cuh"
#include "engine.h"

namespace mcv {

__constant__ int8_t c_pattern[1024] = {
#include "rbrief_pattern.inc"
};

#ifndef MCV_DESC_WARPS
#define MCV_DESC_WARPS 3
#endif
constexpr int DESC_WARPS = MCV_DESC_WARPS;
constexpr int DESC_KPW = 4;      // keypoints per warp and round: one per group of 8 lanes
#ifndef DESC_WAVES
#define DESC_WAVES 4             // persistent grid: about this many waves of resident CTAs (3 per SM), split evenly over the images
#endif
#ifndef DESC_PREFETCH
#define DESC_PREFETCH 0          // cp.async.bulk.prefetch.tensor of the warp's next round: measured slower (0.78 vs 0.58 ms per 384
#endif                           // images) — it doubles the row requests the TMA engine has to generate

// Design. A quarter-warp (8 lanes) owns one keypoint for the scalar part of the work (slot decoding, IC_Angle, cv::fastAtan2,
// sincosf, keypoint assembly), which is therefore paid once per FOUR keypoints of a warp instead of once per keypoint; for the
// descriptor the warp turns to its four keypoints one after the other, lane b computing byte b from the 16 pattern points it
// keeps in registers (the group leaders' cos / sin / window offset arrive by shuffle). Both pixel neighbourhoods of the keypoint
// are fetched by the TMA engine (cp.async.bulk.tensor, one instruction per window, issued by the group's first lane, completion
// on the warp's mbarrier) — one after the other into the SAME shared-memory buffer — from per-level 3-D tensor
// maps (x, y, image) over the pyramid / blurred-pyramid buffers:
//  * blurred window for rBRIEF: the pattern points lie within radius 18.4 of the centre, so every steered sample falls inside
//    37 x 37; box 64 x 37 starting at ((cx - 18) & ~15, cy - 18) — the TMA start coordinate of the byte dimension must be a
//    multiple of 16 (an unaligned one raises "illegal instruction": scripts/exp/tma_probe.cu) — and whatever lies past the
//    row pitch is zero-filled. The 512 single-byte gathers then hit shared-memory banks instead of ~20 different L1 lines
//    per warp instruction.
//  * unblurred 31 x 31 patch for IC_Angle: box 48 x 31 at ((cx - 15) & ~15, cy - 15); lane k funnel-shifts two consecutive
//    words of a row into patch pixels 4k .. 4k+3.
// The kernel is persistent per image (a few CTAs per image, tables built once per CTA): a warp loops over rounds of four
// keypoints; the other resident warps (12 per SM) cover the TMA latency of a round.
// History (ncu, 384 images, 770 k keypoints): warp-per-keypoint with global gathers 0.83 ms (636 warp-instructions per
// keypoint, l1tex 96 %); the same with cp.async staging 0.88 ms (810 per keypoint, short-scoreboard bound); quarter-warp + TMA,
// one CTA per 16 keypoints 0.91 ms (table prologue per CTA, 18 % occupancy); persistent 0.58 ms (424 per keypoint; l1tex 78 %,
// L2 56 %: what moves is ~3.9 KB of window per keypoint, 3 GB per launch); the two windows one after the other through ONE
// buffer per keypoint (2.4 instead of 3.9 KB: 20 resident warps per SM instead of 12) 0.567 -> 0.532 ms; lane per descriptor
// byte with the pattern in registers 0.532 -> 0.498 ms, 3 warps per CTA 0.486 ms.
constexpr int TILE_R = 18, TILE_ROWS = 2 * TILE_R + 1, TILE_PITCH = 64;
constexpr int IC_ROWS = 31, IC_PITCH = 48;
constexpr int TILE_BYTES = TILE_ROWS * TILE_PITCH;                    // 2368
constexpr int IC_BYTES = IC_ROWS * IC_PITCH;                          // 1488
constexpr int BUF_BYTES = (TILE_BYTES + 127) & ~127;                  // per keypoint; TMA destinations are 128-byte aligned. ONE buffer:
                                                                      // the unblurred patch (IC_BYTES) first, then the blurred window over it
static_assert(IC_BYTES <= BUF_BYTES, "the orientation patch shares the descriptor window's buffer");
constexpr int DESC_DYN_SMEM = DESC_WARPS * DESC_KPW * BUF_BYTES;
constexpr int DESC_RESIDENT = (227 * 1024) / (DESC_DYN_SMEM + 6 * 1024);   // CTAs per SM (dynamic windows + ~5 KB of static tables each)

struct DescMaps {                 // per level: (pitch, h, n_images) u8 tensors over the pyramid and the blurred pyramid
    CUtensorMap pyr[MAX_LEVELS];
    CUtensorMap blur[MAX_LEVELS];
};

__device__ __forceinline__ int dp4a_us(unsigned a, unsigned b_signed, int c) {   // sum of u8(a) * s8(b) + c
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_signed), "r"(c));
    return d;
}
__global__ void __launch_bounds__(32 * DESC_WARPS) k_orient_desc(const __grid_constant__ DescMaps maps,
                                                                 const uint32_t* __restrict__ out_pts, const int* __restrict__ out_cnt,
                                                                 const mcv_keypoint* __restrict__ seeds, int n_seeds,
                                                                 mcv_keypoint* __restrict__ kps, uint8_t* __restrict__ desc,
                                                                 int* __restrict__ counts, int cap, const __grid_constant__ Plan P) {
    extern __shared__ __align__(128) uint8_t s_buf[];                   // [warp][group][BUF_BYTES]
    // pattern staging: s_pat[k][q] = point 64 * q + k; read once per CTA — lane b keeps points 16 b .. 16 b + 15 in registers
    __shared__ float2 s_pat[64][8];
    // IC_Angle weights of the 31x31 circular patch, by row (v + 15) and 4-px word k (u = -15 + 4k ...): s_wu = u, s_wv = v
    // inside the circle (|u| <= umax[|v|], ORBextractor.cc:444-456), 0 outside, as signed bytes for dp4a
    __shared__ unsigned s_wu[32][8], s_wv[32][8];
    __shared__ int s_end[MAX_LEVELS + 1], s_nq[MAX_LEVELS], s_ns[MAX_LEVELS];
    __shared__ __align__(8) unsigned long long s_mbar[DESC_WARPS];
    const int img = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_pat[i & 63][i >> 6] = make_float2((float)c_pattern[2 * i], (float)c_pattern[2 * i + 1]);
    for (int e = threadIdx.x; e < 256; e += blockDim.x) {
        constexpr int UMAX[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
        const int row = e >> 3, k = e & 7;
        unsigned wu = 0, wv = 0;
        if (row < 31) {
            const int v = row - 15, d = UMAX[v < 0 ? -v : v];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int u = -15 + 4 * k + j;
                if (u >= -d && u <= d) { wu |= (unsigned)(u & 0xff) << (8 * j); wv |= (unsigned)(v & 0xff) << (8 * j); }
            }
        }
        s_wu[row][k] = wu; s_wv[row][k] = wv;
    }
    // per-level slot ranges of this image: quadtree keypoints of the level, then the caller's seeds of that octave
    if (warp == 0) {
        int my_nq = 0, my_ns = 0;
        if (lane < P.n_levels) {
            my_nq = out_cnt[(size_t)img * P.n_levels + lane];
            for (int s = 0; s < n_seeds; ++s) my_ns += seeds[s].octave == lane;   // n_seeds is 0 on the batch path
        }
        int my_end = my_nq + my_ns;                                     // inclusive prefix = first slot after this level
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, my_end, o); if (lane >= o) my_end += v; }
        if (lane < MAX_LEVELS) { s_end[lane + 1] = my_end; s_nq[lane] = my_nq; s_ns[lane] = my_ns; }
        if (lane == 0) s_end[0] = 0;
    }
    const unsigned mbar = (unsigned)__cvta_generic_to_shared(&s_mbar[warp]);
    if (lane == 0) {
        mbar_init(mbar, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the init is visible to the TMA engine
    }
    __syncthreads();
    float2 pat[16];                                                     // this lane's 16 pattern points (descriptor byte `lane`)
#pragma unroll
    for (int j = 0; j < 16; ++j) pat[j] = s_pat[16 * (lane & 3) + j][lane >> 2];
    const int total = min(s_end[P.n_levels], cap);
    if (blockIdx.x == 0 && threadIdx.x == 0) counts[img] = total;
    const int grp = lane >> 3, l8 = lane & 7;
    const unsigned gmask = 0xffu << (8 * grp);
    uint8_t* buf = s_buf + (size_t)(warp * DESC_KPW + grp) * BUF_BYTES;
    const unsigned buf_s = (unsigned)__cvta_generic_to_shared(buf);
    const int n_rounds = (total + DESC_KPW - 1) / DESC_KPW;
    const int round_stride = gridDim.x * DESC_WARPS;

    // slot -> level, centre. `pt` = packed quadtree point (level coordinates - BORDER); a pre-seeded keypoint is the
    // (j - n_quad)-th seed of its octave in the caller's order (ORBextractor.cc:845-847)
    auto decode = [&](int slot, bool& valid, int& level, int& cx, int& cy, int& seed_idx, uint32_t& pt) {
        valid = slot < total; level = 0; cx = cy = 0; seed_idx = -1; pt = 0;
        if (!valid) return;
        while (level + 1 < P.n_levels && slot >= s_end[level + 1]) ++level;
        const int n_quad = s_nq[level];
        const int j = slot - s_end[level];
        if (j < n_quad) {
            pt = __ldg(out_pts + (size_t)img * P.out_per_image + P.lv[level].out_off + j);
            cx = pt_x(pt) + BORDER; cy = pt_y(pt) + BORDER;
        } else {
            int k = j - n_quad, s = 0;
            for (; s < n_seeds; ++s) if (seeds[s].octave == level && k-- == 0) break;
            seed_idx = s;
            cx = cv_round_f(seeds[s].x); cy = cv_round_f(seeds[s].y);
        }
    };

    int round = blockIdx.x * DESC_WARPS + warp;
    bool valid, valid_n; int level, level_n, cx, cx_n, cy, cy_n, seed_idx, seed_idx_n; uint32_t pt, pt_n;
    decode(round * DESC_KPW + grp, valid, level, cx, cy, seed_idx, pt);
    unsigned parity = 0;
    for (; round < n_rounds; round += round_stride) {                   // warp-uniform
        const int slot = round * DESC_KPW + grp;
        const unsigned n_valid = __popc(__ballot_sync(0xffffffffu, valid)) >> 3;   // groups with a keypoint (>= 1 here)
        // Two passes over the SAME shared-memory buffer (2.4 KB per keypoint instead of 3.9 KB for two: 20 resident warps per SM
        // instead of 12, which is what this latency- and shared-memory-bound kernel wants): the unblurred patch for IC_Angle,
        // then — once the moments are summed — the blurred window for rBRIEF, whose flight covers atan2 / sincosf.
        if (lane == 0) mbar_expect_tx(mbar, n_valid * (unsigned)IC_BYTES);
        __syncwarp();
        if (valid && l8 == 0) tma_load_3d(buf_s, &maps.pyr[level], (cx - 15) & ~15, cy - 15, img, mbar);
        // the warp's next round
        decode(slot + round_stride * DESC_KPW, valid_n, level_n, cx_n, cy_n, seed_idx_n, pt_n);
        if (DESC_PREFETCH && valid_n && l8 == 0) {
            tma_prefetch_3d(&maps.blur[level_n], (cx_n - TILE_R) & ~15, cy_n - TILE_R, img);
            tma_prefetch_3d(&maps.pyr[level_n], (cx_n - 15) & ~15, cy_n - 15, img);
        }
        mbar_wait(mbar, parity);
        parity ^= 1u;
        int m10 = 0, m01 = 0;
        if (valid && seed_idx < 0) {                                    // whole groups; the shuffles below are group-wide
            // IC_Angle: lane l8 owns patch columns 4 l8 .. 4 l8 + 3 of every row; the row's 31 patch bytes start `ox` bytes
            // into the staged row. The four groups walk the rows with an offset of grp so that their loads (buffers are
            // 128-byte aligned) spread over the banks.
            const int ox = (cx - 15) & 15, a8 = (ox & 3) * 8;
            const unsigned* ic = reinterpret_cast<const unsigned*>(buf) + (ox >> 2) + l8;
#pragma unroll
            for (int i = 0; i < IC_ROWS; ++i) {
                int row = i + grp;
                row = row >= IC_ROWS ? row - IC_ROWS : row;
                const unsigned w = __funnelshift_r(ic[row * (IC_PITCH / 4)], ic[row * (IC_PITCH / 4) + 1], a8);
                m10 = dp4a_us(w, s_wu[row][l8], m10);
                m01 = dp4a_us(w, s_wv[row][l8], m01);
            }
        }
        // the patch has been read (generic proxy): the blurred window may land on it (async proxy)
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (lane == 0) mbar_expect_tx(mbar, n_valid * (unsigned)TILE_BYTES);
        __syncwarp();
        if (valid && l8 == 0) tma_load_3d(buf_s, &maps.blur[level], (cx - TILE_R) & ~15, cy - TILE_R, img, mbar);
        mcv_keypoint kp;
        float a = 1.f, bs = 0.f;
        if (valid) {
            if (seed_idx < 0) {
                kp.x = (float)cx; kp.y = (float)cy;
                kp.size = (float)P.lv[level].kp_size; kp.response = (float)pt_r(pt); kp.octave = level; kp.class_id = -1;
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) { m10 += __shfl_xor_sync(gmask, m10, o); m01 += __shfl_xor_sync(gmask, m01, o); }
                kp.angle = fast_atan2_deg((float)m01, (float)m10);
            } else {
                kp = seeds[seed_idx];
            }
            const float ang = __fmul_rn(kp.angle, 0.017453292519943295f);  // factorPI = (float)(CV_PI / 180.f)
            sincosf_glibc(ang, &bs, &a);  // a = cos, bs = sin
        }
        mbar_wait(mbar, parity);
        parity ^= 1u;
        // steered BRIEF: the whole warp on one keypoint at a time, lane b = descriptor byte b (pattern points 16 b .. 16 b + 15 sit
        // in this lane's registers: no pattern loads in the loop; a quarter-warp per keypoint with the pattern in shared memory
        // spent 64 of its ~190 shared-memory instructions per round on them).
        // cvRound by the 1.5 * 2^23 trick (round-half-even like cvRound's lrint; |v| < 19): integer = float bits -
        // 0x4B400000, and that constant, times 65 for row * 64 + column, is folded into the window's base offset (mod 2^32)
        {
            constexpr float MAGIC = 12582912.0f;
            const unsigned base = (unsigned)(TILE_R * TILE_PITCH + TILE_R + ((cx - TILE_R) & 15)) - (unsigned)(TILE_PITCH + 1) * 0x4B400000u;
#pragma unroll
            for (int gg = 0; gg < DESC_KPW; ++gg) {
                if (!__shfl_sync(0xffffffffu, (int)valid, 8 * gg)) continue;           // warp-uniform
                const float ag = __shfl_sync(0xffffffffu, a, 8 * gg), bg = __shfl_sync(0xffffffffu, bs, 8 * gg);
                const unsigned base_g = __shfl_sync(0xffffffffu, base, 8 * gg);
                const uint8_t* wb = s_buf + (size_t)(warp * DESC_KPW + gg) * BUF_BYTES;
                unsigned val = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const float2 q0 = pat[2 * k], q1 = pat[2 * k + 1];
                    const unsigned r0 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(q0.x, bg), __fmul_rn(q0.y, ag)), MAGIC));
                    const unsigned c0 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(q0.x, ag), __fmul_rn(q0.y, bg)), MAGIC));
                    const unsigned r1 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(q1.x, bg), __fmul_rn(q1.y, ag)), MAGIC));
                    const unsigned c1 = __float_as_uint(__fadd_rn(__fsub_rn(__fmul_rn(q1.x, ag), __fmul_rn(q1.y, bg)), MAGIC));
                    const int t0 = wb[r0 * TILE_PITCH + c0 + base_g], t1 = wb[r1 * TILE_PITCH + c1 + base_g];
                    val |= (unsigned)(t0 < t1) << k;
                }
                desc[((size_t)img * cap + (size_t)round * DESC_KPW + gg) * 32 + lane] = (uint8_t)val;
            }
        }
        if (valid) {
            const LevelGeom& g = P.lv[level];
            if (level != 0) { kp.x = __fmul_rn(kp.x, g.scale); kp.y = __fmul_rn(kp.y, g.scale); }
            if (l8 == 0) kps[(size_t)img * cap + slot] = kp;
        }
        // the buffers are about to be overwritten through the async proxy: order the generic-proxy reads before it
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        valid = valid_n; level = level_n; cx = cx_n; cy = cy_n; seed_idx = seed_idx_n; pt = pt_n;
    }
}

int launch_orient_desc(const Plan& P, const uint8_t* d_pyr, const uint8_t* d_blur, const uint32_t* d_out_pts, const int* d_out_cnt,
                       const SeedInfo* seeds, mcv_keypoint* d_kps, uint8_t* d_desc, int* d_counts, int cap, int n_images,
                       cudaStream_t s) {
    const int n_seeds = seeds ? seeds->n_seeds : 0;
    const int max_kp = std::max(1, std::min(cap, P.max_quad_kp + n_seeds));
    constexpr int per_cta = DESC_WARPS * DESC_KPW;
    const int ctas_per_image = std::max(1, std::min((max_kp + per_cta - 1) / per_cta, (DESC_WAVES * DESC_RESIDENT * NUM_SMS + n_images - 1) / n_images));
    dim3 grid(ctas_per_image, n_images);
    DescMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int l = 0; l < P.n_levels; ++l)
        if (!encode_level_map(&maps.pyr[l], d_pyr, P, l, n_images, IC_PITCH, IC_ROWS) ||
            !encode_level_map(&maps.blur[l], d_blur, P, l, n_images, TILE_PITCH, TILE_ROWS)) {
            set_error("cuTensorMapEncodeTiled failed for the orientation / descriptor windows");
            return -1;
        }
    cudaFuncSetAttribute(k_orient_desc, cudaFuncAttributeMaxDynamicSharedMemorySize, DESC_DYN_SMEM);   // per device; cheap
    k_orient_desc<<<grid, 32 * DESC_WARPS, DESC_DYN_SMEM, s>>>(maps, d_out_pts, d_out_cnt, seeds ? seeds->d_seeds : nullptr, n_seeds, d_kps,
                                                              d_desc, d_counts, cap, P);
    return 1;
}

}  // namespace mcv
