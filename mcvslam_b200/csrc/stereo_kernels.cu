// Left/right stereo matching for a batch of frames (sm_100a). Replaces Frame::ComputeStereoMatch (src/Frame.cpp:150-328).
//
// k_stereo_rows: one CTA per frame. Counting sort of the right keypoints by (octave, image row floor(yR)) into a compact record
//     array (uR, row band [minr, maxr], octave, iR) + row_start[octave * h + row] — the device form of vRowIndices
//     (src/Frame.cpp:154-168). A right keypoint of octave o can only be a candidate of left row `row` if
//     |floor(yR) - row| <= ceil(10 * scale[o]) + 1, and only for left keypoints of octaves o - 1 .. o + 1, so the match kernel
//     scans three short slices of the sorted records (~2-3 % of them; a row-only table with the band of the coarsest octave made
//     it 16 %) instead of all right keypoints.
// k_stereo_match: one warp per left keypoint.
//   * candidate gate, evaluated exactly as the reference does on every record of that slice:
//     row band floor(yR - r) <= (int)vL <= ceil(yR + r) with r = 10 * scale[octR]  (src/Frame.cpp:160-168),
//     |octR - octL| <= 1, uL - maxD <= uR <= uL - 1                                 (src/Frame.cpp:199-215);
//     the 2-NN key carries iR, so the result does not depend on the order the slice is scanned in;
//   * Hamming 2-NN over the candidates with strict '<' streaming semantics == lexicographic (distance, index) top-2
//     (Matcher::KnnMatch + LoopBody, src/Matcher.cpp:245-302), sentinel distance 999;
//   * FilterRatio(0.70) and FilterThreshold(int(46 * 0.75) = 34)                    (src/Frame.cpp:225);
//   * 11x11 SAD of centre-subtracted patches at 11 integer shifts on the keypoint's pyramid level, parabola sub-pixel fit,
//     disparity / depth                                                            (src/Frame.cpp:228-303).
// k_stereo_median: one CTA per frame; the (size/2)-th smallest SAD via a two-pass radix select, then invalidates matches
//     with SAD > 1.6 * median or < 0.4 * median                                    (src/Frame.cpp:307-323).
#include <stdlib.h>
#include <algorithm>
#include "devmath.cuh"
#include "engine.h"

namespace mcv {

// The kernel is bound by the latency of its chain of dependent loads (keypoint -> row table -> records -> descriptors -> patches),
// not by issue slots: what helps is more warps in flight. 2 warps per CTA (a CTA holds its slot until its slowest warp is done)
// and 40 registers (48 warps per SM; a few spills outside the loops). B200, 128 frames: 8 warps / 51 registers 0.350 ms,
// 4 / 51: 0.340, 2 / 51: 0.326, 8 / 40: 0.298, 4 / 40: 0.280, 2 / 40: 0.275, 4 / 32: 0.313. Issuing the left patch loads ahead of the
// candidate scan and taking the winner's uR from its lane instead of kr[] (one round trip less) was bit-exact and no faster.
#ifndef MCV_ST_WARPS
#define MCV_ST_WARPS 2
#endif
#ifndef MCV_ST_MINB
#define MCV_ST_MINB 24
#endif
constexpr int ST_WARPS = MCV_ST_WARPS;

struct StereoArgs {
    const uint8_t* pyr_l; const uint8_t* pyr_r;    // image-0 pyramid bases; frame f adds f * frame_pyr_stride
    size_t frame_pyr_stride;
    const mcv_keypoint* kl; const mcv_keypoint* kr; const uint8_t* dl; const uint8_t* dr;
    const int* nl; const int* nr;                  // per-frame counts (stride count_stride), or NULL -> nl_fixed / nr_fixed
    int nl_fixed, nr_fixed;
    size_t frame_kp_stride;                        // keypoints between consecutive frames (kl/kr/dl/dr/outputs of L)
    int count_stride;
    float bf, baseline;
    float* u_right; float* depth; int* best_dist; int* best_r;
    size_t out_stride;
    int* row_start;                                // [frame][n_levels * hb + 2]: first sorted record of each (octave, row bin) (+ end)
    int row_shift, hb;                             // row bin = row >> row_shift, hb bins per octave (shift 0 unless the table would
                                                   // not fit the rows kernel's shared memory: many levels of a very tall image)
    uint4* recs;                                   // [frame][rec_stride] records sorted by (octave, row)
    size_t rec_stride;
};

// ---------------------------------------------------------------------------------------------------------
// row table of the right keypoints: record = (uR bits, minr | maxr << 16, iR | octave << 24, row)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_stereo_rows(const __grid_constant__ StereoArgs A, const __grid_constant__ Plan P) {
    extern __shared__ int s_row[];                 // n_levels * h + 1 counters, then reused as running offsets
    const int f = blockIdx.x, tid = threadIdx.x, h = P.h, bins = P.n_levels * A.hb;
    const int nr = A.nr ? A.nr[(size_t)f * A.count_stride] : A.nr_fixed;
    const mcv_keypoint* kr = A.kr + (size_t)f * A.frame_kp_stride;
    int* row_start = A.row_start + (size_t)f * (bins + 2);
    uint4* recs = A.recs + (size_t)f * A.rec_stride;
    auto bin_of = [&](const mcv_keypoint& k) {
        return min(max(k.octave, 0), P.n_levels - 1) * A.hb + (min(max((int)floorf(k.y), 0), h - 1) >> A.row_shift);
    };
    for (int i = tid; i <= bins; i += 256) s_row[i] = 0;
    __syncthreads();
    for (int i = tid; i < nr; i += 256) atomicAdd(&s_row[bin_of(kr[i])], 1);
    __syncthreads();
    // exclusive scan over the bins: 256 threads x ceil(bins / 256) consecutive bins each, warp shuffles + one smem hop
    __shared__ int s_warp[8];
    const int per = (bins + 255) / 256, r0 = tid * per;
    int sum = 0;
    for (int r = r0; r < min(r0 + per, bins); ++r) sum += s_row[r];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
    if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
    __syncthreads();
    int base = incl - sum;
    for (int w = 0; w < (tid >> 5); ++w) base += s_warp[w];
    for (int r = r0; r < min(r0 + per, bins); ++r) { const int c = s_row[r]; s_row[r] = base; row_start[r] = base; base += c; }
    if (tid == 255) { row_start[bins] = nr; row_start[bins + 1] = nr; }
    __syncthreads();
    for (int i = tid; i < nr; i += 256) {
        const float uR = kr[i].x, yR = kr[i].y;
        const int oct = kr[i].octave;
        const float r = __fmul_rn(10.f, P.lv[oct].scale);
        const int maxr = (int)ceilf(__fadd_rn(yR, r)), minr = (int)floorf(__fsub_rn(yR, r));
        const int pos = atomicAdd(&s_row[bin_of(kr[i])], 1);
        // rows are clamped to 16 bits: |minr|, |maxr| < 32768 for any supported image (<= 4128 rows)
        recs[pos] = make_uint4(__float_as_uint(uR), (unsigned)(minr & 0xffff) | ((unsigned)maxr << 16), (unsigned)i | ((unsigned)oct << 24), 0u);
    }
}

__global__ void __launch_bounds__(32 * ST_WARPS, MCV_ST_MINB) k_stereo_match(const __grid_constant__ StereoArgs A, const __grid_constant__ Plan P) {
    const int f = blockIdx.y, lane = threadIdx.x & 31;
    const int iL = blockIdx.x * ST_WARPS + (threadIdx.x >> 5);
    const int nl = A.nl ? A.nl[(size_t)f * A.count_stride] : A.nl_fixed;
    if (iL >= nl) return;
    const mcv_keypoint* kl = A.kl + (size_t)f * A.frame_kp_stride;
    const mcv_keypoint* kr = A.kr + (size_t)f * A.frame_kp_stride;   // .x of the matched keypoint only
    const uint8_t* dl = A.dl + (size_t)f * A.frame_kp_stride * 32;
    const uint8_t* dr = A.dr + (size_t)f * A.frame_kp_stride * 32;
    float* o_ur = A.u_right + (size_t)f * A.out_stride;
    float* o_dp = A.depth + (size_t)f * A.out_stride;
    int* o_bd = A.best_dist + (size_t)f * A.out_stride;
    int* o_br = A.best_r ? A.best_r + (size_t)f * A.out_stride : nullptr;
    if (lane == 0) { o_ur[iL] = -1.0f; o_dp[iL] = -1.0f; o_bd[iL] = -1; if (o_br) o_br[iL] = -1; }

    const int n_rows = P.h;
    const float uL = kl[iL].x, vL = kl[iL].y;
    const int levelL = kl[iL].octave;
    const float maxD = fminf(__fdiv_rn(A.bf, A.baseline), 1000.0f);
    const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, 1.0f);
    if (vL < 0.f || (int)vL >= n_rows || maxU < 0.f) return;
    const int row = (int)vL;
    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(dl + (size_t)iL * 32));
    const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(dl + (size_t)iL * 32 + 16));
    // per-lane streaming top-2 over this lane's candidates (ascending iR), keys = dist << 20 | iR
    const unsigned SENT = 999u << 20;
    unsigned k0 = SENT, k1 = SENT;
    {
        const int* row_start = A.row_start + (size_t)f * (P.n_levels * A.hb + 2);
        const uint4* recs = A.recs + (size_t)f * A.rec_stride;
        // the slices of octaves levelL - 1 .. levelL + 1: rows row -+ (ceil(10 * scale[o]) + 2) of each, one concatenated index space
        int beg[3], cum[3];
        int total = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int o = levelL - 1 + k;
            beg[k] = 0;
            if (o >= 0 && o < P.n_levels) {
                const int b = (int)ceilf(__fmul_rn(10.f, P.lv[o].scale)) + 2;
                beg[k] = __ldg(row_start + o * A.hb + (max(row - b, 0) >> A.row_shift));
                total += __ldg(row_start + o * A.hb + (min(row + b, n_rows - 1) >> A.row_shift) + 1) - beg[k];
            }
            cum[k] = total;
        }
        for (int p = lane; p < total; p += 32) {
            const int idx = p < cum[0] ? beg[0] + p : (p < cum[1] ? beg[1] + (p - cum[0]) : beg[2] + (p - cum[1]));
            const uint4 rec = __ldg(recs + idx);
            const float uR = __uint_as_float(rec.x);
            const int minr = (int)(short)(rec.y & 0xffffu), maxr = (int)rec.y >> 16;
            const int octR = (int)(rec.z >> 24), iR = (int)(rec.z & 0xffffffu);
            if (row >= minr && row <= maxr && octR >= levelL - 1 && octR <= levelL + 1 && uR >= minU && uR <= maxU) {
                const uint4 t0 = __ldg(reinterpret_cast<const uint4*>(dr + (size_t)iR * 32));
                const uint4 t1 = __ldg(reinterpret_cast<const uint4*>(dr + (size_t)iR * 32 + 16));
                const unsigned key = ((unsigned)hamming256(q0, q1, t0, t1) << 20) | (unsigned)iR;
                if (key < k0) { k1 = k0; k0 = key; } else if (key < k1) k1 = key;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned o0 = __shfl_xor_sync(0xffffffffu, k0, o), o1 = __shfl_xor_sync(0xffffffffu, k1, o);
        const unsigned n0 = min(k0, o0), n1 = min(max(k0, o0), min(k1, o1));
        k0 = n0; k1 = n1;
    }
    if (k0 == SENT) return;  // no candidate
    const int d0 = (int)(k0 >> 20), d1 = (int)(k1 >> 20), bestR = (int)(k0 & 0xfffffu);
    if (!(__fdiv_rn((float)d0, (float)d1) <= 0.70f)) return;   // FilterRatio(0.70)
    if ((float)d0 > 34.0f) return;                             // FilterThreshold(46 * 0.75 -> int 34)

    // sub-pixel refinement on level `levelL` of both pyramids
    const LevelGeom& g = P.lv[levelL];
    const float uR0 = kr[bestR].x;
    const float sc = g.inv_scale;
    const float suL = roundf(__fmul_rn(uL, sc)), svL = roundf(__fmul_rn(vL, sc)), suR0 = roundf(__fmul_rn(uR0, sc));
    const int w = 5, L = 5;
    if (svL - w < 0 || svL + w + 1 >= g.h || suL - w < 0 || suL + w + 1 >= g.w) return;
    if (suR0 + L - w < 0 || suR0 + L + w + 1 >= g.w) return;
    if (suR0 - L - w < 0) return;  // the reference would throw on this negative colRange; unreachable for quadtree keypoints
    const uint8_t* IL = A.pyr_l + (size_t)f * A.frame_pyr_stride + g.img_off;
    const uint8_t* IR = A.pyr_r + (size_t)f * A.frame_pyr_stride + g.img_off;
    const int cy = (int)svL, cxL = (int)suL, cxR = (int)suR0;
    const int lc = IL[(size_t)cy * g.pitch + cxL];
    int sad[11];
#pragma unroll
    for (int s = 0; s < 11; ++s) sad[s] = 0;
    const uint8_t* rrow_c = IR + (size_t)cy * g.pitch + cxR;
    for (int p = lane; p < 121; p += 32) {
        const int yy = p / 11 - w, xx = p % 11 - w;
        const int lv = IL[(size_t)(cy + yy) * g.pitch + cxL + xx] - lc;
        const uint8_t* rr = IR + (size_t)(cy + yy) * g.pitch + cxR + xx;
#pragma unroll
        for (int s = 0; s < 11; ++s) sad[s] += abs(lv - ((int)rr[s - L] - (int)rrow_c[s - L]));
    }
    // warp sums, two shifts per register (a full SAD is at most 121 * 510 = 61710 < 2^16: no carry between the halves)
#pragma unroll
    for (int s = 0; s < 11; s += 2) {
        unsigned pr = (unsigned)sad[s] | (s + 1 < 11 ? (unsigned)sad[s + 1] << 16 : 0u);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) pr += __shfl_xor_sync(0xffffffffu, pr, o);
        sad[s] = (int)(pr & 0xffffu);
        if (s + 1 < 11) sad[s + 1] = (int)(pr >> 16);
    }
    int bestDist = 0x7fffffff, bestinc = 0;
#pragma unroll
    for (int s = 0; s < 11; ++s) if (sad[s] < bestDist) { bestDist = sad[s]; bestinc = s - L; }
    if (bestinc == -L || bestinc == L) return;
    float dist1 = 0.f, dist2 = 0.f, dist3 = 0.f;
#pragma unroll
    for (int s = 1; s < 10; ++s) if (s - L == bestinc) { dist1 = (float)sad[s - 1]; dist2 = (float)sad[s]; dist3 = (float)sad[s + 1]; }
    const float deltaR = __fdiv_rn(__fsub_rn(dist1, dist3), __fmul_rn(2.0f, __fsub_rn(__fadd_rn(dist1, dist3), __fmul_rn(2.0f, dist2))));
    if (deltaR < -1.f || deltaR > 1.f) return;
    const float bestuR = __fmul_rn(g.scale, __fadd_rn(__fadd_rn(suR0, (float)bestinc), deltaR));
    const float disparity = __fsub_rn(uL, bestuR);
    if (disparity >= 1.0f && disparity < maxD) {
        if (lane == 0) {
            o_dp[iL] = __fdiv_rn(A.bf, disparity);
            o_ur[iL] = bestuR;
            o_bd[iL] = bestDist;
            if (o_br) o_br[iL] = bestR;
        }
    }
}

// SAD values are < 2^16 (121 * 510 = 61710).
__global__ void __launch_bounds__(256) k_stereo_median(const __grid_constant__ StereoArgs A) {
    __shared__ int hist[256];
    __shared__ int s_sel, s_rem, s_total;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int nl = A.nl ? A.nl[(size_t)f * A.count_stride] : A.nl_fixed;
    float* o_ur = A.u_right + (size_t)f * A.out_stride;
    float* o_dp = A.depth + (size_t)f * A.out_stride;
    const int* o_bd = A.best_dist + (size_t)f * A.out_stride;
    hist[tid] = 0;
    if (tid == 0) s_total = 0;
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < nl; i += 256) { const int d = o_bd[i]; if (d >= 0) { atomicAdd(&hist[d >> 8], 1); ++mine; } }
    if (mine) atomicAdd(&s_total, mine);
    __syncthreads();
    const int total = s_total;
    if (total == 0) return;
    if (tid == 0) {
        int k = total / 2, b = 0;  // index (size_t)(size * 1.0 / 2) of the sorted list
        while (k >= hist[b]) { k -= hist[b]; ++b; }
        s_sel = b; s_rem = k;
    }
    __syncthreads();
    const int hi = s_sel, k2 = s_rem;
    __syncthreads();
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < nl; i += 256) { const int d = o_bd[i]; if (d >= 0 && (d >> 8) == hi) atomicAdd(&hist[d & 255], 1); }
    __syncthreads();
    if (tid == 0) {
        int k = k2, b = 0;
        while (k >= hist[b]) { k -= hist[b]; ++b; }
        s_sel = (hi << 8) | b;
    }
    __syncthreads();
    const float median = (float)s_sel;
    const float th_max = __fmul_rn(1.6f, median);
    const float th_min = (float)(0.4 * (double)median);
    for (int i = tid; i < nl; i += 256) {
        const int d = o_bd[i];
        if (d >= 0 && ((float)d > th_max || (float)d < th_min)) { o_ur[i] = -1.0f; o_dp[i] = -1.0f; }
    }
}

size_t stereo_scratch_bytes(const Plan& P, int n_frames, int max_right) {
    return (size_t)n_frames * (((size_t)P.n_levels * P.h + 2) * sizeof(int) + (size_t)max_right * sizeof(uint4)) + 256;
}

static int run_stereo(StereoArgs A, const Plan& P, int n_frames, int max_left, int max_right, void* scratch, cudaStream_t s, cudaEvent_t mid = nullptr) {
    // scratch layout: records first (16-byte aligned), then the row tables
    A.recs = reinterpret_cast<uint4*>(scratch);
    A.rec_stride = (size_t)max_right;
    A.row_start = reinterpret_cast<int*>(A.recs + (size_t)n_frames * max_right);
    // one bin per (octave, row) while that table fits 64 KB of the rows kernel's shared memory; coarser row bins beyond (the match
    // kernel's gate is exact either way, it only scans a few more records). MCV_STEREO_ROW_SHIFT forces a shift (tests).
    A.row_shift = 0;
    while (((size_t)P.n_levels * (((P.h - 1) >> A.row_shift) + 1) + 1) * sizeof(int) > 64 * 1024) ++A.row_shift;
    if (const char* e = getenv("MCV_STEREO_ROW_SHIFT")) A.row_shift = std::max(A.row_shift, std::min(8, atoi(e)));
    A.hb = ((P.h - 1) >> A.row_shift) + 1;
    const size_t smem = ((size_t)P.n_levels * A.hb + 1) * sizeof(int);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_stereo_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_stereo_rows<<<n_frames, 256, smem, s>>>(A, P);
    dim3 grid((max_left + ST_WARPS - 1) / ST_WARPS, n_frames);
    k_stereo_match<<<grid, 32 * ST_WARPS, 0, s>>>(A, P);
    if (mid) cudaEventRecord(mid, s);
    k_stereo_median<<<n_frames, 256, 0, s>>>(A);
    return 3;
}

int launch_stereo(const Plan& P, const uint8_t* d_pyr, const mcv_keypoint* d_kps, const uint8_t* d_desc, const int* d_counts, int cap,
                  int n_frames, int left_cam, int right_cam, int cams_per_frame, float bf, float baseline, float* d_u_right, float* d_depth,
                  int* d_best_dist, int* d_best_r, void* d_scratch, cudaStream_t s, cudaEvent_t mid) {
    StereoArgs A{};
    A.pyr_l = d_pyr + (size_t)left_cam * P.pyr_bytes; A.pyr_r = d_pyr + (size_t)right_cam * P.pyr_bytes;
    A.frame_pyr_stride = (size_t)cams_per_frame * P.pyr_bytes;
    A.kl = d_kps + (size_t)left_cam * cap; A.kr = d_kps + (size_t)right_cam * cap;
    A.dl = d_desc + (size_t)left_cam * cap * 32; A.dr = d_desc + (size_t)right_cam * cap * 32;
    A.nl = d_counts + left_cam; A.nr = d_counts + right_cam; A.count_stride = cams_per_frame;
    A.frame_kp_stride = (size_t)cams_per_frame * cap;
    A.bf = bf; A.baseline = baseline;
    A.u_right = d_u_right; A.depth = d_depth; A.best_dist = d_best_dist; A.best_r = d_best_r; A.out_stride = cap;
    return run_stereo(A, P, n_frames, cap, cap, d_scratch, s, mid);
}

// Rig outputs are fixed-capacity records (cap slots per image); the slots behind an image's count are defined too: zero
// keypoints / descriptors, -1 ("none", as Frame::ComputeStereoMatch initialises them, src/Frame.cpp:152-153) for u_right / depth.
// One CTA per image; a frame's u_right / depth rows are padded by the CTA of its left image.
__global__ void __launch_bounds__(128) k_fill_tails(mcv_keypoint* __restrict__ kps, uint8_t* __restrict__ desc, const int* __restrict__ counts, int cap,
                                                    float* __restrict__ u_right, float* __restrict__ depth, int cams_per_frame) {
    const int img = blockIdx.x, n = min(counts[img], cap);
    uint32_t* k = reinterpret_cast<uint32_t*>(kps + (size_t)img * cap + n);            // 7 words per keypoint
    for (int i = threadIdx.x; i < (cap - n) * 7; i += blockDim.x) k[i] = 0u;
    uint32_t* d = reinterpret_cast<uint32_t*>(desc + ((size_t)img * cap + n) * 32);    // 8 words per descriptor
    for (int i = threadIdx.x; i < (cap - n) * 8; i += blockDim.x) d[i] = 0u;
    if (img % cams_per_frame == 0 && u_right) {
        const size_t f = (size_t)(img / cams_per_frame) * cap;
        for (int i = n + threadIdx.x; i < cap; i += blockDim.x) { u_right[f + i] = -1.f; depth[f + i] = -1.f; }
    }
}

int launch_fill_tails(mcv_keypoint* d_kps, uint8_t* d_desc, const int* d_counts, int cap, int n_images, float* d_u_right, float* d_depth,
                      int cams_per_frame, cudaStream_t s) {
    if (n_images <= 0) return 0;
    k_fill_tails<<<n_images, 128, 0, s>>>(d_kps, d_desc, d_counts, cap, d_u_right, d_depth, cams_per_frame);
    return 1;
}

int launch_stereo_pair(const Plan& P, const uint8_t* d_pyr_l, const uint8_t* d_pyr_r, const mcv_keypoint* d_kl, const uint8_t* d_dl, int nl,
                       const mcv_keypoint* d_kr, const uint8_t* d_dr, int nr, float bf, float baseline, float* d_u_right, float* d_depth,
                       int* d_best_dist, int* d_best_r, void* d_scratch, cudaStream_t s) {
    StereoArgs A{};
    A.pyr_l = d_pyr_l; A.pyr_r = d_pyr_r; A.frame_pyr_stride = 0;
    A.kl = d_kl; A.kr = d_kr; A.dl = d_dl; A.dr = d_dr;
    A.nl = nullptr; A.nr = nullptr; A.nl_fixed = nl; A.nr_fixed = nr; A.frame_kp_stride = 0; A.count_stride = 0;
    A.bf = bf; A.baseline = baseline;
    A.u_right = d_u_right; A.depth = d_depth; A.best_dist = d_best_dist; A.best_r = d_best_r; A.out_stride = 0;
    if (nl <= 0) return 0;
    return run_stereo(A, P, 1, nl, std::max(nr, 1), d_scratch, s);
}

}  // namespace mcv
