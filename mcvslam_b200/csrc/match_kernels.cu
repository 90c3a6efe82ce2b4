// Hamming 2-NN matching kernels (sm_100a). Integer pipe only: LOP3 (xor), POPC, IADD3, IMNMX.
//
//   k_knn2_bf          : brute force, replaces cv::BFMatcher(NORM_HAMMING).knnMatch as used by Matcher::KnnMatch(Mat, Mat)
//                        (src/Matcher.cpp:304-308) and the O(Q*T) LoopBody (src/Matcher.cpp:245-281). Result per query =
//                        the two smallest (distance, trainIdx) pairs in lexicographic order, which is what both the
//                        OpenCV matcher and the first-party strict-'<' streaming loop produce.
//   k_knn2_candidates  : the same 2-NN over a per-query candidate list (src/Frame.cpp:199-225, src/Object.cpp:217-226,
//                        src/Map.cpp:495-527, src/Matcher.cpp:162-181); ties resolve by position in the list.
//   k_project_match    : Object::ProjectBunchMapPoints (src/Object.cpp:208-236) incl. the 30x30 grid window query
//                        (src/Object.cpp:249-308) and FilterRatio(0.6) / FilterThreshold(46).
//
// A (distance, index) pair is packed into one unsigned key, distance in the high bits, so a lexicographic top-2 update is
// three integer min/max instructions and a cross-lane / cross-CTA merge is the same operation.
#include "devmath.cuh"
#include "engine.h"

namespace mcv {

constexpr int BF_THREADS = 128;        // queries per CTA (one per thread)
constexpr int BF_TILE = 256;           // train descriptors staged in shared memory per step
constexpr int BF_IDX_BITS = 22;        // trainIdx bits in the key; distance (<= 256) above them
constexpr unsigned BF_SENT = 0xffffffffu;

__device__ __forceinline__ void top2_update(unsigned& k0, unsigned& k1, unsigned key) {
    k1 = min(k1, max(k0, key));
    k0 = min(k0, key);
}

__device__ __forceinline__ void warp_top2_merge(unsigned& k0, unsigned& k1) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned o0 = __shfl_xor_sync(0xffffffffu, k0, o), o1 = __shfl_xor_sync(0xffffffffu, k1, o);
        const unsigned n0 = min(k0, o0), n1 = min(max(k0, o0), min(k1, o1));
        k0 = n0; k1 = n1;
    }
}

// grid = (ceil(nq / BF_THREADS), n_splits). Each CTA scans train rows [split * per_split, ...) for its queries and writes
// the partial top-2 keys to part[split][q][2].
__global__ void __launch_bounds__(BF_THREADS) k_knn2_bf(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t, int nt,
                                                        int per_split, unsigned* __restrict__ part) {
    __shared__ uint4 s_t[BF_TILE * 2];
    const int qi = blockIdx.x * BF_THREADS + threadIdx.x;
    const int t_begin = blockIdx.y * per_split, t_end = min(nt, t_begin + per_split);
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (qi < nq) {
        a0 = __ldg(reinterpret_cast<const uint4*>(q + (size_t)qi * 32));
        a1 = __ldg(reinterpret_cast<const uint4*>(q + (size_t)qi * 32 + 16));
    }
    unsigned k0 = BF_SENT, k1 = BF_SENT;
    for (int base = t_begin; base < t_end; base += BF_TILE) {
        const int n = min(BF_TILE, t_end - base);
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * n; i += BF_THREADS) s_t[i] = __ldg(reinterpret_cast<const uint4*>(t + (size_t)base * 32) + i);
        __syncthreads();
        unsigned key_base = (unsigned)base;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const uint4 b0 = s_t[2 * j], b1 = s_t[2 * j + 1];
            const unsigned d = (unsigned)hamming256(a0, a1, b0, b1);
            top2_update(k0, k1, (d << BF_IDX_BITS) + key_base + (unsigned)j);
        }
    }
    if (qi < nq) {
        unsigned* o = part + ((size_t)blockIdx.y * nq + qi) * 2;
        o[0] = k0; o[1] = k1;
    }
}

__global__ void __launch_bounds__(256) k_knn2_merge(const unsigned* __restrict__ part, int nq, int n_splits, int train_offset,
                                                    int32_t* __restrict__ idx, int32_t* __restrict__ dist) {
    const int qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= nq) return;
    unsigned k0 = BF_SENT, k1 = BF_SENT;
    for (int s = 0; s < n_splits; ++s) {
        const unsigned* p = part + ((size_t)s * nq + qi) * 2;
        top2_update(k0, k1, p[0]);
        top2_update(k0, k1, p[1]);
    }
    const unsigned mask = (1u << BF_IDX_BITS) - 1;
    idx[2 * qi] = k0 == BF_SENT ? -1 : (int)(k0 & mask) + train_offset;
    dist[2 * qi] = k0 == BF_SENT ? 0x7fffffff : (int)(k0 >> BF_IDX_BITS);
    idx[2 * qi + 1] = k1 == BF_SENT ? -1 : (int)(k1 & mask) + train_offset;
    dist[2 * qi + 1] = k1 == BF_SENT ? 0x7fffffff : (int)(k1 >> BF_IDX_BITS);
}

// Mid-sized problems (one image against another: 2000 x 2000) in ONE launch: a warp per query, 16 queries per CTA, the train set
// streamed through shared memory 256 rows at a time (the next tile's global load is in flight while the current one is scanned),
// lane l takes rows l, l + 32, ...; the warp's 32 partial top-2 pairs are merged by shuffles and lane 0 writes the result. With
// k_knn2_bf such a call is 128 CTAs of one warp per scheduler plus a merge launch (22.7 us at 2000 x 2000, 31 % of the POPC peak).
constexpr int WQ_WARPS = 16;
constexpr int WQ_TILE = 256;

__global__ void __launch_bounds__(32 * WQ_WARPS) k_knn2_wq(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t, int nt,
                                                           int train_offset, int32_t* __restrict__ idx, int32_t* __restrict__ dist) {
    __shared__ uint4 s_lo[WQ_TILE], s_hi[WQ_TILE];         // the two halves of a row apart: consecutive rows are 16 bytes apart
    const int tid = threadIdx.x, lane = tid & 31;
    const int qi = blockIdx.x * WQ_WARPS + (tid >> 5);
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (qi < nq) {
        a0 = __ldg(reinterpret_cast<const uint4*>(q + (size_t)qi * 32));
        a1 = __ldg(reinterpret_cast<const uint4*>(q + (size_t)qi * 32 + 16));
    }
    const uint4* t4 = reinterpret_cast<const uint4*>(t);
    unsigned k0 = BF_SENT, k1 = BF_SENT;
    uint4 pre = tid < 2 * min(WQ_TILE, nt) ? __ldg(t4 + tid) : make_uint4(0, 0, 0, 0);
    for (int base = 0; base < nt; base += WQ_TILE) {
        const int n = min(WQ_TILE, nt - base);
        __syncthreads();                                   // the previous tile has been scanned
        ((tid & 1) ? s_hi : s_lo)[tid >> 1] = pre;
        __syncthreads();
        const int nb = base + WQ_TILE;
        if (nb < nt && tid < 2 * min(WQ_TILE, nt - nb)) pre = __ldg(t4 + (size_t)nb * 2 + tid);
#pragma unroll 4
        for (int j = lane; j < n; j += 32) {
            const unsigned d = (unsigned)hamming256(a0, a1, s_lo[j], s_hi[j]);
            top2_update(k0, k1, (d << BF_IDX_BITS) + (unsigned)(base + j));
        }
    }
    warp_top2_merge(k0, k1);
    if (lane == 0 && qi < nq) {
        const unsigned mask = (1u << BF_IDX_BITS) - 1;
        idx[2 * qi] = k0 == BF_SENT ? -1 : (int)(k0 & mask) + train_offset;
        dist[2 * qi] = k0 == BF_SENT ? 0x7fffffff : (int)(k0 >> BF_IDX_BITS);
        idx[2 * qi + 1] = k1 == BF_SENT ? -1 : (int)(k1 & mask) + train_offset;
        dist[2 * qi + 1] = k1 == BF_SENT ? 0x7fffffff : (int)(k1 >> BF_IDX_BITS);
    }
}

// the one-launch kernel pays when a warp's share of the train set is short (measured against the split + merge pair:
// scripts/r03/match_small.py); from 2^23 pairs up the tensor-core path takes over
static bool wq_usable(int nq, int nt) {
    if (getenv("MCV_KNN_NO_WQ")) return false;             // comparison runs only (scripts/r03/match_small.py); read per call
    return nq >= 16 * WQ_WARPS && nt <= 4096 && (long long)nq * nt < (1 << 23);   // its time grows with nt alone: ~4 us + 5.2 ns per train row
}

// Split geometry of one brute-force call: enough CTAs for ~4 per SM, but at least BF_TILE train rows each.
static void bf_splits(int nq, int nt, int& q_blocks, int& n_splits, int& per_split) {
    q_blocks = (nq + BF_THREADS - 1) / BF_THREADS;
    n_splits = std::max(1, std::min((4 * NUM_SMS + q_blocks - 1) / q_blocks, (nt + BF_TILE - 1) / BF_TILE));
    per_split = nt > 0 ? ((nt + n_splits - 1) / n_splits + BF_TILE - 1) / BF_TILE * BF_TILE : BF_TILE;
    n_splits = nt > 0 ? (nt + per_split - 1) / per_split : 1;
}

// MCV_KNN_POPC=1 keeps every brute-force call on the integer-pipe kernel (comparison runs, ncu of k_knn2_bf)
static bool force_popc() {
    const char* e = getenv("MCV_KNN_POPC");   // read per call: bench.py flips it between its two matching legs
    return e && atoi(e) != 0;
}

// bytes of scratch one call needs; the CALLER owns it (per thread / per stream), the launchers keep no state
size_t knn2_bf_part_bytes(int nq, int nt) {
    if (nq <= 0) return 0;
    if (!force_popc() && knn2_tc_usable(nq, nt)) return knn2_tc_scratch_bytes(nq, nt, nq, nt, 1);
    if (wq_usable(nq, nt)) return 0;                       // one launch, no partial keys
    int q_blocks, n_splits, per_split;
    bf_splits(nq, nt, q_blocks, n_splits, per_split);
    return (size_t)n_splits * nq * 2 * sizeof(unsigned);
}

// Brute-force 2-NN: tensor-core path (match_tc_kernels.cu) from 2^23 pairs up, integer-pipe kernel below that.
int launch_knn2_bf(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int train_offset, int32_t* d_idx, int32_t* d_dist, unsigned* d_part,
                   cudaStream_t s) {
    if (nq <= 0) return 0;
    if (nt > (1 << BF_IDX_BITS)) return -1;
    const bool tc = !force_popc() && knn2_tc_usable(nq, nt);
    if (!d_part && (tc || !wq_usable(nq, nt))) return -1;
    if (tc) return launch_knn2_tc(d_q, nq, d_t, nt, train_offset, d_idx, d_dist, d_part, 0, nullptr, nullptr, nullptr, 1, s);
    if (wq_usable(nq, nt)) {
        k_knn2_wq<<<(nq + WQ_WARPS - 1) / WQ_WARPS, 32 * WQ_WARPS, 0, s>>>(d_q, nq, d_t, nt, train_offset, d_idx, d_dist);
        return 1;
    }
    int q_blocks, n_splits, per_split;
    bf_splits(nq, nt, q_blocks, n_splits, per_split);
    k_knn2_bf<<<dim3(q_blocks, n_splits), BF_THREADS, 0, s>>>(d_q, nq, d_t, nt, per_split, d_part);
    k_knn2_merge<<<(nq + 255) / 256, 256, 0, s>>>(d_part, nq, n_splits, train_offset, d_idx, d_dist);
    return 2;
}

// ---------------------------------------------------------------------------------------------------------
// candidate lists: one warp per query
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_knn2_candidates(const uint8_t* __restrict__ q, int nq, const uint8_t* __restrict__ t,
                                                         const int32_t* __restrict__ off, const int32_t* __restrict__ cidx,
                                                         int32_t* __restrict__ idx, int32_t* __restrict__ dist) {
    const int qi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (qi >= nq) return;
    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(q + (size_t)qi * 32));
    const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(q + (size_t)qi * 32 + 16));
    const int b = off[qi], e = off[qi + 1];
    const unsigned SENT = 999u << 20;  // the reference's d = {999, 999}, idx = {0, 0} (src/Matcher.cpp:258-259)
    unsigned k0 = SENT, k1 = SENT;
    for (int p = b + lane; p < e; p += 32) {
        const int j = cidx[p];
        const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(t + (size_t)j * 32));
        const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(t + (size_t)j * 32 + 16));
        top2_update(k0, k1, ((unsigned)hamming256(a0, a1, b0, b1) << 20) | (unsigned)(p - b));
    }
    warp_top2_merge(k0, k1);
    if (lane == 0) {
        idx[2 * qi] = (int)(k0 & 0xfffffu); dist[2 * qi] = (int)(k0 >> 20);
        idx[2 * qi + 1] = (int)(k1 & 0xfffffu); dist[2 * qi + 1] = (int)(k1 >> 20);
    }
}

int launch_knn2_candidates(const uint8_t* d_q, int nq, const uint8_t* d_t, const int32_t* d_off, const int32_t* d_cidx, int32_t* d_idx,
                           int32_t* d_dist, cudaStream_t s) {
    if (nq <= 0) return 0;
    k_knn2_candidates<<<(nq + 7) / 8, 256, 0, s>>>(d_q, nq, d_t, d_off, d_cidx, d_idx, d_dist);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------
// Object::AssignFeaturesToGrid / PosInGrid (src/Object.cpp:182-201,249-257) on the device: the 30 x 30 cell table of one image's
// keypoints in the order GetFeaturesInArea walks it (cell-column-major = key gx * 30 + gy, ascending keypoint index inside a
// cell). cell_start[901] = exclusive prefix of the cell populations, cell_idx[n] = keypoint indices. One CTA: shared-memory
// histogram, scan, unordered scatter, then one thread per cell puts its (short) list in ascending order.
// ---------------------------------------------------------------------------------------------------------
constexpr int GRID_N = 30;  // FRAME_GRID_COLS / ROWS, include/Object.hpp:27-28
constexpr int GRID_CELLS = GRID_N * GRID_N;

__device__ __forceinline__ int grid_cell_of(float kx, float ky, float winv, float hinv) {
    // PosInGrid: round(), keypoints whose cell falls outside [0, 30) are not in the grid at all
    const int gx = (int)roundf(__fmul_rn(kx, winv)), gy = (int)roundf(__fmul_rn(ky, hinv));
    return (gx < 0 || gx >= GRID_N || gy < 0 || gy >= GRID_N) ? -1 : gx * GRID_N + gy;
}

__global__ void __launch_bounds__(1024) k_grid_build(const mcv_keypoint* __restrict__ kps, int n, int w, int h, int32_t* __restrict__ cell_start,
                                                     int32_t* __restrict__ cell_idx) {
    __shared__ int s_cnt[GRID_CELLS], s_start[GRID_CELLS + 1];
    const int tid = threadIdx.x;
    const float winv = (float)((double)GRID_N / w), hinv = (float)((double)GRID_N / h);
    for (int c = tid; c < GRID_CELLS; c += blockDim.x) s_cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) { const int c = grid_cell_of(kps[i].x, kps[i].y, winv, hinv); if (c >= 0) atomicAdd(&s_cnt[c], 1); }
    __syncthreads();
    if (tid < 32) {   // exclusive scan of 900 counters by one warp (29 per lane)
        const int per = (GRID_CELLS + 31) / 32, lo = min(tid * per, GRID_CELLS), hi = min(lo + per, GRID_CELLS);
        int sum = 0;
        for (int c = lo; c < hi; ++c) sum += s_cnt[c];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += t; }
        int run = incl - sum;
        for (int c = lo; c < hi; ++c) { s_start[c] = run; run += s_cnt[c]; }
        if (tid == 31) s_start[GRID_CELLS] = run;
    }
    __syncthreads();
    for (int c = tid; c < GRID_CELLS; c += blockDim.x) s_cnt[c] = s_start[c];   // scatter cursors
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) { const int c = grid_cell_of(kps[i].x, kps[i].y, winv, hinv); if (c >= 0) cell_idx[atomicAdd(&s_cnt[c], 1)] = i; }
    __syncthreads();
    for (int c = tid; c < GRID_CELLS; c += blockDim.x) {   // ascending keypoint index inside the cell (insertion sort, lists are short)
        const int lo = s_start[c], hi = s_start[c + 1];
        for (int a = lo + 1; a < hi; ++a) {
            const int v = cell_idx[a];
            int b = a;
            while (b > lo && cell_idx[b - 1] > v) { cell_idx[b] = cell_idx[b - 1]; --b; }
            cell_idx[b] = v;
        }
    }
    for (int c = tid; c <= GRID_CELLS; c += blockDim.x) cell_start[c] = s_start[c];
}

int launch_grid_build(const mcv_keypoint* d_kps, int n, int w, int h, int32_t* d_cell_start, int32_t* d_cell_idx, cudaStream_t s) {
    k_grid_build<<<1, 1024, 0, s>>>(d_kps, n, w, h, d_cell_start, d_cell_idx);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------
// projection matching: one warp per MapPoint over the cell table's window (the columns of cells [min_cx, max_cx] x [min_cy, max_cy]
// are contiguous runs of cell_idx) with a 64-bit key (distance, grid-cell-major candidate order), so ties resolve exactly like
// the reference's candidate list.
// pose = Rcw (9, row-major) | tcw (3) | fx fy cx cy
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_project_match(const mcv_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, int n, int w,
                                                       int h, const float* __restrict__ scale, const float* __restrict__ pose,
                                                       const float* __restrict__ xyz, const uint8_t* __restrict__ mp_desc,
                                                       const int32_t* __restrict__ mp_level, int n_mp, float r_th,
                                                       const int32_t* __restrict__ cell_start, const int32_t* __restrict__ cell_idx,
                                                       int32_t* __restrict__ out_idx, int32_t* __restrict__ out_dist) {
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= n_mp) return;
    if (lane == 0) { out_idx[m] = -1; out_dist[m] = -1; }
    const float X = xyz[3 * m], Y = xyz[3 * m + 1], Z = xyz[3 * m + 2];
    float pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)  // cv::Mat 3x3 * 3x1 + 3x1 in float, left to right (small-matrix gemm path)
        pc[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(pose[3 * r], X), __fmul_rn(pose[3 * r + 1], Y)), __fmul_rn(pose[3 * r + 2], Z)), pose[9 + r]);
    if (pc[2] < 0.f) return;
    // Pinhole::project (modules/camera/Pinhole.cpp:45-47)
    const float x = __fadd_rn(__fdiv_rn(__fmul_rn(pose[12], pc[0]), pc[2]), pose[14]);
    const float y = __fadd_rn(__fdiv_rn(__fmul_rn(pose[13], pc[1]), pc[2]), pose[15]);
    const float r = __fmul_rn(r_th, scale[mp_level[m]]);
    // GetFeaturesInArea (src/Object.cpp:263-308)
    if (!(x >= 0.f && y >= 0.f && x < (float)w && y < (float)h)) return;
    const float winv = (float)((double)GRID_N / w), hinv = (float)((double)GRID_N / h);
    const int min_cx = max(0, (int)floorf(__fmul_rn(__fsub_rn(x, r), winv)));
    const int max_cx = min(GRID_N - 1, (int)ceilf(__fmul_rn(__fadd_rn(x, r), winv)));
    const int min_cy = max(0, (int)floorf(__fmul_rn(__fsub_rn(y, r), hinv)));
    const int max_cy = min(GRID_N - 1, (int)ceilf(__fmul_rn(__fadd_rn(y, r), hinv)));
    if (min_cx >= GRID_N || max_cx < 0 || min_cy >= GRID_N || max_cy < 0) return;
    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(mp_desc + (size_t)m * 32));
    const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(mp_desc + (size_t)m * 32 + 16));
    const unsigned long long SENT = 999ull << 32;
    unsigned long long k0 = SENT, k1 = SENT;
    for (int gx = min_cx; gx <= max_cx; ++gx) {
        const int pb = cell_start[gx * GRID_N + min_cy], pe = cell_start[gx * GRID_N + max_cy + 1];
        for (int p = pb + lane; p < pe; p += 32) {
            const int j = cell_idx[p];
            const float kx = kps[j].x, ky = kps[j].y;
            if (!(fabsf(__fsub_rn(kx, x)) < r && fabsf(__fsub_rn(ky, y)) < r)) continue;
            const int gy = (int)roundf(__fmul_rn(ky, hinv));
            const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)j * 32));
            const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)j * 32 + 16));
            // candidate order: ix outer, iy inner, ascending keypoint index inside a cell
            const unsigned long long key = ((unsigned long long)hamming256(a0, a1, b0, b1) << 32) | ((unsigned long long)(gx * GRID_N + gy) << 21) | (unsigned)j;
            const unsigned long long hi = key > k0 ? key : k0;
            k1 = k1 < hi ? k1 : hi;
            k0 = k0 < key ? k0 : key;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, k0, o), o1 = __shfl_xor_sync(0xffffffffu, k1, o);
        const unsigned long long lo = k0 < o0 ? k0 : o0, hi = k0 < o0 ? o0 : k0, l1 = k1 < o1 ? k1 : o1;
        k0 = lo; k1 = hi < l1 ? hi : l1;
    }
    if (k0 == SENT) return;
    const float d0 = (float)(unsigned)(k0 >> 32), d1 = (float)(unsigned)(k1 >> 32);
    if (!(__fdiv_rn(d0, d1) <= 0.6f)) return;  // FilterRatio() default
    if (d0 > 46.0f) return;                    // FilterThreshold() default ORB_GOOD_THRESHOLD
    if (lane == 0) { out_idx[m] = (int)(k0 & 0x1fffffu); out_dist[m] = (int)(k0 >> 32); }
}

int launch_project(const mcv_keypoint* d_kps, const uint8_t* d_desc, int n, int w, int h, const float* d_scale, const float* d_pose,
                   const float* d_xyz, const uint8_t* d_mp_desc, const int32_t* d_level, int n_mp, float r_th, int32_t* d_cell_start,
                   int32_t* d_cell_idx, int32_t* d_idx, int32_t* d_dist, cudaStream_t s) {
    if (n_mp <= 0) return 0;
    launch_grid_build(d_kps, n, w, h, d_cell_start, d_cell_idx, s);
    k_project_match<<<(n_mp + 7) / 8, 256, 0, s>>>(d_kps, d_desc, n, w, h, d_scale, d_pose, d_xyz, d_mp_desc, d_level, n_mp, r_th, d_cell_start, d_cell_idx,
                                                   d_idx, d_dist);
    return 2;
}

// ---------------------------------------------------------------------------------------------------------
// Window matching shared by Map::Fuse and Tracker::Wnd_Track: GetFeaturesInArea(x, y, r) (src/Object.cpp:263-308) over the
// 30x30 grid of AssignFeaturesToGrid (src/Object.cpp:182-201,249-257) + first-party 2-NN (src/Matcher.cpp:256-275). One warp
// per query; lanes stride over the keypoints; a candidate's key is (distance, grid-cell-major order, index), so the packed
// minimum reproduces the reference's candidate-list order for ties. `gate(j)` is the caller's extra per-candidate test.
// Returns in every lane the two smallest keys (SENT when absent) and, in `first`, the smallest order key (first candidate).
// ---------------------------------------------------------------------------------------------------------
constexpr unsigned long long WM_SENT = 999ull << 32;

template <typename Gate>
__device__ __forceinline__ void window_knn2(const mcv_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, int n, int w, int h, float x,
                                            float y, float r, uint4 a0, uint4 a1, Gate gate, unsigned long long& k0, unsigned long long& k1,
                                            unsigned& first, const int32_t* __restrict__ cell_start, const int32_t* __restrict__ cell_idx) {
    const int lane = threadIdx.x & 31;
    k0 = WM_SENT; k1 = WM_SENT; first = 0xffffffffu;
    bool ok = x >= 0.f && y >= 0.f && x < (float)w && y < (float)h;
    const float winv = (float)((double)GRID_N / w), hinv = (float)((double)GRID_N / h);
    const int min_cx = max(0, (int)floorf(__fmul_rn(__fsub_rn(x, r), winv)));
    const int max_cx = min(GRID_N - 1, (int)ceilf(__fmul_rn(__fadd_rn(x, r), winv)));
    const int min_cy = max(0, (int)floorf(__fmul_rn(__fsub_rn(y, r), hinv)));
    const int max_cy = min(GRID_N - 1, (int)ceilf(__fmul_rn(__fadd_rn(y, r), hinv)));
    ok = ok && !(min_cx >= GRID_N || max_cx < 0 || min_cy >= GRID_N || max_cy < 0);
    if (ok) {
        for (int gx = min_cx; gx <= max_cx; ++gx) {
            const int pb = cell_start[gx * GRID_N + min_cy], pe = cell_start[gx * GRID_N + max_cy + 1];
            for (int p = pb + lane; p < pe; p += 32) {
                const int j = cell_idx[p];
                const float kx = kps[j].x, ky = kps[j].y;
                if (!(fabsf(__fsub_rn(kx, x)) < r && fabsf(__fsub_rn(ky, y)) < r)) continue;
                if (!gate(j)) continue;
                const int gy = (int)roundf(__fmul_rn(ky, hinv));
                const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)j * 32));
                const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)j * 32 + 16));
                const unsigned order = ((unsigned)(gx * GRID_N + gy) << 21) | (unsigned)j;
                const unsigned long long key = ((unsigned long long)hamming256(a0, a1, b0, b1) << 32) | order;
                const unsigned long long hi = key > k0 ? key : k0;
                k1 = k1 < hi ? k1 : hi;
                k0 = k0 < key ? k0 : key;
                first = min(first, order);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long o0 = __shfl_xor_sync(0xffffffffu, k0, o), o1 = __shfl_xor_sync(0xffffffffu, k1, o);
        const unsigned long long lo = k0 < o0 ? k0 : o0, hi = k0 < o0 ? o0 : k0, l1 = k1 < o1 ? k1 : o1;
        k0 = lo; k1 = hi < l1 ? hi : l1;
        first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
    }
}

// FilterRatio() / FilterThreshold() defaults on one 2-NN row (src/Matcher.cpp:100-111,22-33; include/Matcher.hpp:14,55)
__device__ __forceinline__ bool ratio_threshold_pass(unsigned long long k0, unsigned long long k1) {
    if (k0 == WM_SENT) return false;
    const float d0 = (float)(unsigned)(k0 >> 32), d1 = (float)(unsigned)(k1 >> 32);
    return __fdiv_rn(d0, d1) <= 0.6f && !(d0 > 46.0f);
}

// Map::Fuse matching front-end (src/Map.cpp:478-527). par = Rcw (9) | tcw (3) | fx fy cx cy | Ow (3) | bf | sigma2[L] | inv_sigma2[L]
__global__ void __launch_bounds__(256) k_fuse_match(const mcv_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc, int n, int w, int h,
                                                    const float* __restrict__ par, int n_levels, const float* __restrict__ depth_left,
                                                    const float* __restrict__ xyz, const float* __restrict__ normal,
                                                    const uint8_t* __restrict__ mp_desc, const int32_t* __restrict__ mp_level, int n_mp,
                                                    const int32_t* __restrict__ cell_start, const int32_t* __restrict__ cell_idx,
                                                    int32_t* __restrict__ out_idx, int32_t* __restrict__ out_dist) {
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (m >= n_mp) return;                                   // warp-uniform
    if (lane == 0) { out_idx[m] = -1; out_dist[m] = -1; }
    const float X = xyz[3 * m], Y = xyz[3 * m + 1], Z = xyz[3 * m + 2];
    // viewing angle (:483-488): PO = Xw - Ow in float; cv::norm and Mat::dot accumulate the products in double
    const float po0 = __fsub_rn(X, par[16]), po1 = __fsub_rn(Y, par[17]), po2 = __fsub_rn(Z, par[18]);
    const double n2 = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)po0), __dmul_rn((double)po1, (double)po1)), __dmul_rn((double)po2, (double)po2));
    const float dist3d = __double2float_rn(__dsqrt_rn(n2));
    const double dot = __dadd_rn(__dadd_rn(__dmul_rn((double)po0, (double)normal[3 * m]), __dmul_rn((double)po1, (double)normal[3 * m + 1])),
                                 __dmul_rn((double)po2, (double)normal[3 * m + 2]));
    if (dot < __dmul_rn(0.5, (double)dist3d)) return;
    float pc[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)  // Object::Map: mRcw * x3D + mtcw, float gemm left to right
        pc[r] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(par[3 * r], X), __fmul_rn(par[3 * r + 1], Y)), __fmul_rn(par[3 * r + 2], Z)), par[9 + r]);
    const float z = pc[2];
    if (z <= 0.f) return;
    const float invz = __double2float_rn(__ddiv_rn(1.0, (double)z));   // const float invz = 1. / z
    const float u = __fadd_rn(__fdiv_rn(__fmul_rn(par[12], pc[0]), pc[2]), par[14]);
    const float v = __fadd_rn(__fdiv_rn(__fmul_rn(par[13], pc[1]), pc[2]), par[15]);
    const float ur = __fsub_rn(u, __fmul_rn(par[19], invz));
    const int lvl = mp_level[m], lvl_lo = max(0, lvl - 1);
    const float* sigma2 = par + 20;
    const float* inv_sigma2 = par + 20 + n_levels;
    auto gate = [&](int j) {
        const int level = kps[j].octave;
        if (level < lvl_lo || level > lvl) return false;
        const float ex = __fsub_rn(u, kps[j].x), ey = __fsub_rn(v, kps[j].y);
        const float dl = depth_left[j];
        if (dl >= 0.f) {   // "stereo" branch exactly as written (:505-516)
            const float er = __fsub_rn(ur, dl);
            const float e2 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(er, er));
            return !((double)__fmul_rn(e2, sigma2[level]) > 7.8);
        }
        const float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
        return !((double)__fmul_rn(e2, inv_sigma2[level]) > 5.99);
    };
    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(mp_desc + (size_t)m * 32));
    const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(mp_desc + (size_t)m * 32 + 16));
    unsigned long long k0, k1; unsigned first;
    window_knn2(kps, desc, n, w, h, u, v, 10.f, a0, a1, gate, k0, k1, first, cell_start, cell_idx);
    if (!ratio_threshold_pass(k0, k1)) return;
    if (lane == 0) { out_idx[m] = (int)(k0 & 0x1fffffu); out_dist[m] = (int)(k0 >> 32); }
}

// Tracker::Wnd_Track (src/Tracker.cpp:341-360): +-20 px window of obj2 around every listed keypoint of obj1. out_idx = the
// index the reference reports (candi_idxs[queryIdx] = the window's FIRST candidate), out_best = the matched candidate.
__global__ void __launch_bounds__(256) k_wnd_track(const mcv_keypoint* __restrict__ kps1, const uint8_t* __restrict__ desc1,
                                                   const int32_t* __restrict__ q_idx, int n_q, const mcv_keypoint* __restrict__ kps2,
                                                   const uint8_t* __restrict__ desc2, int n2, int w, int h,
                                                   const int32_t* __restrict__ cell_start, const int32_t* __restrict__ cell_idx, int32_t* __restrict__ out_idx,
                                                   int32_t* __restrict__ out_best, int32_t* __restrict__ out_dist) {
    const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= n_q) return;
    if (lane == 0) { out_idx[q] = -1; out_best[q] = -1; out_dist[q] = -1; }
    const int i = q_idx[q];
    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(desc1 + (size_t)i * 32));
    const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(desc1 + (size_t)i * 32 + 16));
    unsigned long long k0, k1; unsigned first;
    window_knn2(kps2, desc2, n2, w, h, kps1[i].x, kps1[i].y, 20.f, a0, a1, [](int) { return true; }, k0, k1, first, cell_start, cell_idx);
    if (!ratio_threshold_pass(k0, k1)) return;
    if (lane == 0) { out_idx[q] = (int)(first & 0x1fffffu); out_best[q] = (int)(k0 & 0x1fffffu); out_dist[q] = (int)(k0 >> 32); }
}

int launch_fuse_match(const mcv_keypoint* d_kps, const uint8_t* d_desc, int n, int w, int h, const float* d_par, int n_levels,
                      const float* d_depth_left, const float* d_xyz, const float* d_normal, const uint8_t* d_mp_desc, const int32_t* d_level,
                      int n_mp, int32_t* d_cell_start, int32_t* d_cell_idx, int32_t* d_idx, int32_t* d_dist, cudaStream_t s) {
    if (n_mp <= 0) return 0;
    launch_grid_build(d_kps, n, w, h, d_cell_start, d_cell_idx, s);
    k_fuse_match<<<(n_mp + 7) / 8, 256, 0, s>>>(d_kps, d_desc, n, w, h, d_par, n_levels, d_depth_left, d_xyz, d_normal, d_mp_desc, d_level, n_mp,
                                               d_cell_start, d_cell_idx, d_idx, d_dist);
    return 2;
}

int launch_wnd_track(const mcv_keypoint* d_kps1, const uint8_t* d_desc1, const int32_t* d_qidx, int n_q, const mcv_keypoint* d_kps2,
                     const uint8_t* d_desc2, int n2, int w, int h, int32_t* d_cell_start, int32_t* d_cell_idx, int32_t* d_idx, int32_t* d_best,
                     int32_t* d_dist, cudaStream_t s) {
    if (n_q <= 0) return 0;
    launch_grid_build(d_kps2, n2, w, h, d_cell_start, d_cell_idx, s);
    k_wnd_track<<<(n_q + 7) / 8, 256, 0, s>>>(d_kps1, d_desc1, d_qidx, n_q, d_kps2, d_desc2, n2, w, h, d_cell_start, d_cell_idx, d_idx, d_best, d_dist);
    return 2;
}

// ---------------------------------------------------------------------------------------------------------
// DBoW3::Vocabulary::transform's tree descent (modules/DBow3/src/Vocabulary.cpp:641-672) for a batch of features: one warp per
// feature, lanes = children of the current node (k = 10 in orbvoc). The reference keeps the FIRST child with the smallest
// distance (strict '<'), so the key is (distance, position in the children list). out_leaf = the leaf node reached,
// out_nid = the node passed at depth nid_level = L - levelsup (0 = root when nid_level <= 0).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bow_descend(const uint8_t* __restrict__ desc, int n, const int32_t* __restrict__ child_off,
                                                     const uint32_t* __restrict__ child_ids, const uint8_t* __restrict__ node_desc,
                                                     int nid_level, int max_depth, uint32_t* __restrict__ out_leaf, uint32_t* __restrict__ out_nid) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)i * 32));
    const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)i * 32 + 16));
    unsigned node = 0, nid = 0;
    for (int level = 1; level <= max_depth; ++level) {      // max_depth bounds a malformed (cyclic) child table
        const int lo = child_off[node], hi = child_off[node + 1];
        if (lo == hi) break;                                 // isLeaf()
        unsigned best = 0xffffffffu;
        for (int c = lo + lane; c < hi; c += 32) {
            const unsigned id = child_ids[c];
            const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(node_desc + (size_t)id * 32));
            const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(node_desc + (size_t)id * 32 + 16));
            best = min(best, ((unsigned)hamming256(a0, a1, b0, b1) << 20) | (unsigned)(c - lo));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        node = child_ids[lo + (int)(best & 0xfffffu)];
        if (level == nid_level) nid = node;
    }
    if (lane == 0) { out_leaf[i] = node; out_nid[i] = nid; }
}

int launch_bow_descend(const uint8_t* d_desc, int n, const int32_t* d_child_off, const uint32_t* d_child_ids, const uint8_t* d_node_desc,
                       int nid_level, int max_depth, uint32_t* d_leaf, uint32_t* d_nid, cudaStream_t s) {
    if (n <= 0) return 0;
    k_bow_descend<<<(n + 7) / 8, 256, 0, s>>>(d_desc, n, d_child_off, d_child_ids, d_node_desc, nid_level, max_depth, d_leaf, d_nid);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cpp:101-150) for a batch of MapPoints: the observed descriptors of
// point m are rows [off[m], off[m + 1]) of `desc`. Per point: all-pairs Hamming distances (diagonal 0), every row sorted, its
// median = row[(size_t)(0.5 * (N - 1))], and the FIRST row with the smallest median wins (strict '<' against INT_MAX).
// One CTA per point, one warp per row: distances go into a 257-bin shared-memory histogram (a sorted row is only ever read at
// one rank, so counting replaces std::sort), the rank is located with a warp scan over 9-bin slices, and the point's winner is
// the minimum of the packed key median << 22 | row.
// ---------------------------------------------------------------------------------------------------------
constexpr int DD_WARPS = 8;
constexpr int DD_BINS = 288;                     // 257 distances rounded up to 32 lanes x 9 bins

__global__ void __launch_bounds__(32 * DD_WARPS) k_distinctive(const uint8_t* __restrict__ desc, const int32_t* __restrict__ off, int n_mp,
                                                                 int32_t* __restrict__ best_idx, int32_t* __restrict__ best_median) {
    __shared__ unsigned s_hist[DD_WARPS][DD_BINS];
    __shared__ unsigned s_best;
    const int m = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (m >= n_mp) return;
    const int lo = off[m], N = off[m + 1] - lo;
    if (threadIdx.x == 0) s_best = 0xffffffffu;
    __syncthreads();
    if (N > 0) {
        const int k = (N - 1) >> 1;              // (size_t)(0.5 * (N - 1))
        unsigned* hist = s_hist[warp];
        unsigned mine = 0xffffffffu;
        for (int i = warp; i < N; i += DD_WARPS) {
#pragma unroll
            for (int b = 0; b < DD_BINS / 32; ++b) hist[b * 32 + lane] = 0u;
            __syncwarp();
            const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(lo + i) * 32));
            const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(lo + i) * 32 + 16));
            for (int j = lane; j < N; j += 32) {
                const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(lo + j) * 32));
                const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(desc + (size_t)(lo + j) * 32 + 16));
                atomicAdd(&hist[j == i ? 0 : hamming256(a0, a1, b0, b1)], 1u);   // the buffer's diagonal stays 0 (src/MapPoint.cpp:123)
            }
            __syncwarp();
            unsigned c[9], sum = 0;
#pragma unroll
            for (int b = 0; b < 9; ++b) { c[b] = hist[lane * 9 + b]; sum += c[b]; }
            unsigned incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            // the lane whose slice holds rank k: exclusive prefix <= k < inclusive prefix
            unsigned before = incl - sum;
            int med = -1;
            if (before <= (unsigned)k && (unsigned)k < incl) {
#pragma unroll
                for (int b = 0; b < 9; ++b) { if (med < 0 && before + c[b] > (unsigned)k) med = lane * 9 + b; before += c[b]; }
            }
            const unsigned who = __ballot_sync(0xffffffffu, med >= 0);
            med = __shfl_sync(0xffffffffu, med, __ffs(who) - 1);
            mine = min(mine, ((unsigned)med << 22) | (unsigned)i);
            __syncwarp();
        }
        if (lane == 0) atomicMin(&s_best, mine);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        best_idx[m] = N > 0 ? (int)(s_best & 0x3fffffu) : -1;
        best_median[m] = N > 0 ? (int)(s_best >> 22) : -1;
    }
}

int launch_distinctive(const uint8_t* d_desc, const int32_t* d_off, int n_mp, int32_t* d_best_idx, int32_t* d_best_median, cudaStream_t s) {
    if (n_mp <= 0) return 0;
    k_distinctive<<<n_mp, 32 * DD_WARPS, 0, s>>>(d_desc, d_off, n_mp, d_best_idx, d_best_median);
    return 1;
}

// ---------------------------------------------------------------------------------------------------------
// test taps + integer-pipe peak
// ---------------------------------------------------------------------------------------------------------
__global__ void k_debug_sincosf(const float* a, int n, float* s, float* c) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sincosf_glibc(a[i], &s[i], &c[i]);
}
__global__ void k_debug_atan2(const float* y, const float* x, int n, float* o) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) o[i] = fast_atan2_deg(y[i], x[i]);
}
int launch_debug_sincosf(const float* d_a, int n, float* d_s, float* d_c, cudaStream_t s) {
    k_debug_sincosf<<<(n + 255) / 256, 256, 0, s>>>(d_a, n, d_s, d_c);
    return 1;
}
int launch_debug_atan2(const float* d_y, const float* d_x, int n, float* d_o, cudaStream_t s) {
    k_debug_atan2<<<(n + 255) / 256, 256, 0, s>>>(d_y, d_x, n, d_o);
    return 1;
}

// Sustained xor+popc throughput: 8 independent (xor, popc, add) chains per thread, `iters` rounds.
__global__ void __launch_bounds__(256) k_popc_peak(int iters, unsigned* sink) {
    unsigned a[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * 2654435761u + k * 40503u + blockIdx.x; acc[k] = 0; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[k] += __popc(a[k] ^ (unsigned)i); a[k] += acc[k]; }
    }
    unsigned r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r += acc[k];
    if (r == 0x12345678u) sink[0] = r;
}
int launch_popc_peak(int iters, unsigned* d_sink, int blocks, int threads, cudaStream_t s) {
    k_popc_peak<<<blocks, threads, 0, s>>>(iters, d_sink);
    return 1;
}

}  // namespace mcv
