// Quadtree quota distribution for a batch of pyramid levels (sm_100a).
//
// Replaces ORBextractor::DistributeOctTree + ExtractorNode::DivideNode (ORBextractor.cc:469-580). The reference keeps the
// nodes in a std::priority_queue keyed on the node's point count ONLY, so which of several equally-populated nodes is split
// next — and the order the surviving keypoints come out in — is decided by libstdc++'s push_heap/pop_heap sift order.
// That order is part of the result (keypoint order is API: stereo and projection results are indexed by it), so this
// kernel replays bits/stl_heap.h __push_heap/__adjust_heap step for step.
//
// One warp owns one (image, level) task. Lane 0 drives the heap in shared memory; the whole warp does the data-parallel
// parts: the scan over the level's cell counts, the gather of the cell lists into reference order, the stable 4-way
// partition of a node's points (warp ballots), and the final first-max-response search. Point lists ping-pong between two
// arenas: a node occupies the same [start, start+count) range in either arena and its children are written to the other
// one, so no allocation is needed. The arenas live in SHARED memory whenever the level's candidates fit (pts_cap, sized
// at 5 x quota — several times what textured scenes produce); the ~150 dependent partition passes of a level then run at
// shared-memory latency. Levels with more candidates than that use the global arenas: same code, same result.
#include "engine.h"
#include <stdlib.h>

#include "octree_core.cuh"

namespace mcv {

constexpr int OCT_WARPS_MAX = 4;

struct NodeRec {
    short ulx, uly, urx, bry;  // UL.x, UL.y, UR.x, BR.y (BL/BR.x follow)
    uint32_t start;            // bit 31 = arena (0: A, 1: B)
    uint32_t cnt;
};

struct OctShared {  // per-warp views into dynamic shared memory
    unsigned long long* heap;  // (count << 32) | node id
    NodeRec* nodes;
};

// ---- libstdc++ heap replay (comparator: a.count < b.count) ----
__device__ __forceinline__ uint32_t hcnt(unsigned long long k) { return (uint32_t)(k >> 32); }

__device__ __forceinline__ void heap_push(unsigned long long* h, int& size, unsigned long long value) {
    int hole = size++;
    int parent = (hole - 1) / 2;
    while (hole > 0 && hcnt(h[parent]) < hcnt(value)) {
        h[hole] = h[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h[hole] = value;
}

// std::pop_heap followed by pop_back; the popped element is left at h[size] (as std::pop_heap does) and returned.
__device__ __forceinline__ unsigned long long heap_pop(unsigned long long* h, int& size) {
    const unsigned long long top = h[0];
    if (size > 1) {
        const int len = size - 1;
        const unsigned long long value = h[len];
        h[len] = top;
        int hole = 0, child = 0;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            // both children in ONE 16-byte load: h points at slot 1 of a 16-byte aligned array, so the pair (child - 1, child) =
            // (odd, even) index is 16-byte aligned and the sift-down has a single shared-memory latency per level
            const ulonglong2 pair = *reinterpret_cast<const ulonglong2*>(h + child - 1);
            unsigned long long v = pair.y;
            if (hcnt(pair.y) < hcnt(pair.x)) { child--; v = pair.x; }
            h[hole] = v;
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            h[hole] = h[child - 1];
            hole = child - 1;
        }
        int parent = (hole - 1) / 2;
        while (hole > 0 && hcnt(h[parent]) < hcnt(value)) {
            h[hole] = h[parent];
            hole = parent;
            parent = (hole - 1) / 2;
        }
        h[hole] = value;
    }
    --size;
    return top;
}

__device__ __forceinline__ int quadrant(uint32_t p, int mx, int my) {
    // DivideNode membership (ORBextractor.cc:504-515): x < n1.UR.x ? (y < n1.BR.y ? n1 : n3) : (y < n1.BR.y ? n2 : n4)
    const bool left = pt_x(p) < mx, top = pt_y(p) < my;
    return left ? (top ? 0 : 2) : (top ? 1 : 3);
}

// Stable 4-way partition of src[start, start+cnt) into dst at the same range; returns the four counts in c[].
__device__ __forceinline__ void split_points(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int start, int cnt, int mx, int my,
                                             int lane, int c[4]) {
    const unsigned lt = (1u << lane) - 1;
    if (cnt <= 32) {
        const bool valid = lane < cnt;
        const uint32_t p = valid ? src[start + lane] : 0;
        const int q = valid ? quadrant(p, mx, my) : -1;
        unsigned m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { m[k] = __ballot_sync(0xffffffffu, q == k); c[k] = __popc(m[k]); }
        if (valid) {
            const int base = q == 0 ? 0 : q == 1 ? c[0] : q == 2 ? c[0] + c[1] : c[0] + c[1] + c[2];
            const unsigned mm = q == 0 ? m[0] : q == 1 ? m[1] : q == 2 ? m[2] : m[3];
            dst[start + base + __popc(mm & lt)] = p;
        }
        return;
    }
    c[0] = c[1] = c[2] = c[3] = 0;
    for (int i0 = 0; i0 < cnt; i0 += 32) {
        const int i = i0 + lane;
        const int q = i < cnt ? quadrant(src[start + i], mx, my) : -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) c[k] += __popc(__ballot_sync(0xffffffffu, q == k));
    }
    int run[4] = {0, c[0], c[0] + c[1], c[0] + c[1] + c[2]};
    for (int i0 = 0; i0 < cnt; i0 += 32) {
        const int i = i0 + lane;
        const bool valid = i < cnt;
        const uint32_t p = valid ? src[start + i] : 0;
        const int q = valid ? quadrant(p, mx, my) : -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned m = __ballot_sync(0xffffffffu, q == k);
            if (q == k) dst[start + run[k] + __popc(m & lt)] = p;
            run[k] += __popc(m);
        }
    }
}

// Phase clocks of task 0 (level 0 of image 0) of the last k_octree launch: start, gathered, roots, split loop, drain, selection.
__device__ long long g_oct_clk[8];
#define OCT_CLK(i) do { if (clk && lane == 0) clk[i] = clock64(); } while (0)

// The distribution proper, on points already laid out in arena A [0, M) in reference order.
// Writes the selected points (heap-pop order) to out[0..n) and returns n in every lane.
__device__ int distribute_warp(uint32_t* arena_a, uint32_t* arena_b, int M, int box_w, int box_h, int n_ini, float h_x, int N, int out_cap,
                               unsigned long long* heap, NodeRec* nodes, uint32_t* out, int lane, long long* clk = nullptr) {
    if (M == 0) return 0;
    int heap_size = 0, n_nodes = 0;
    // roots (ORBextractor.cc:533-555): bucket by (int)(x / hX), drop empty roots
    if (n_ini == 1) {
        if (lane == 0) {
            nodes[0] = NodeRec{0, 0, (short)(int)(h_x * 1.0f), (short)box_h, 0u, (uint32_t)M};
            heap_push(heap, heap_size, ((unsigned long long)M << 32) | 0ull);
        }
        n_nodes = 1;
    } else {
        int start = 0;
        const unsigned lt = (1u << lane) - 1;
        for (int r = 0; r < n_ini; ++r) {
            int cnt = 0;
            for (int i0 = 0; i0 < M; i0 += 32) {
                const int i = i0 + lane;
                const uint32_t p = i < M ? arena_a[i] : 0;
                const bool in = i < M && (int)(size_t)__fdiv_rn((float)pt_x(p), h_x) == r;
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (in) arena_b[start + cnt + __popc(m & lt)] = p;
                cnt += __popc(m);
            }
            if (cnt > 0) {
                if (lane == 0) {
                    nodes[n_nodes] = NodeRec{(short)(int)__fmul_rn(h_x, (float)r), 0, (short)(int)__fmul_rn(h_x, (float)(r + 1)), (short)box_h,
                                             (uint32_t)start | 0x80000000u, (uint32_t)cnt};
                    heap_push(heap, heap_size, ((unsigned long long)cnt << 32) | (unsigned long long)n_nodes);
                }
                ++n_nodes;
            }
            start += cnt;
        }
    }
    heap_size = __shfl_sync(0xffffffffu, heap_size, 0);
    __syncwarp();
    OCT_CLK(2);
    // split loop (ORBextractor.cc:557-565). Every iteration grows the heap or halves a box, so 16 * N + 256 iterations are
    // never reached on valid input (unique integer points); the guard only keeps corrupt input from spinning forever, which
    // is what the reference would do.
    for (int guard = 16 * N + 256; heap_size < N && guard > 0; --guard) {
        unsigned long long top = 0;
        if (lane == 0) {
            top = heap[0];
            if (hcnt(top) != 1) heap_pop(heap, heap_size);
        }
        top = __shfl_sync(0xffffffffu, top, 0);
        if (hcnt(top) == 1) break;
        const int id = (int)(uint32_t)top;
        __syncwarp();
        const NodeRec nd = nodes[id];
        // DivideNode (ORBextractor.cc:469-522): halves are ceil((float)extent / 2)
        const int half_x = (nd.urx - nd.ulx + 1) >> 1, half_y = (nd.bry - nd.uly + 1) >> 1;
        const int mx = nd.ulx + half_x, my = nd.uly + half_y;
        const int start = (int)(nd.start & 0x7fffffffu);
        const bool in_b = nd.start >> 31;
        int c[4];
        split_points(in_b ? arena_b : arena_a, in_b ? arena_a : arena_b, start, (int)nd.cnt, mx, my, lane, c);
        __syncwarp();
        if (lane == 0) {
            const uint32_t side = in_b ? 0u : 0x80000000u;
            const short bx[4][4] = {{nd.ulx, nd.uly, (short)mx, (short)my}, {(short)mx, nd.uly, nd.urx, (short)my},
                                    {nd.ulx, (short)my, (short)mx, nd.bry}, {(short)mx, (short)my, nd.urx, nd.bry}};
            int off = start;
            bool reuse = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (c[k] > 0) {
                    const int nid = reuse ? id : n_nodes++;
                    reuse = false;
                    nodes[nid] = NodeRec{bx[k][0], bx[k][1], bx[k][2], bx[k][3], (uint32_t)off | side, (uint32_t)c[k]};
                    heap_push(heap, heap_size, ((unsigned long long)c[k] << 32) | (unsigned long long)nid);
                }
                off += c[k];
            }
        }
        heap_size = __shfl_sync(0xffffffffu, heap_size, 0);
        n_nodes = __shfl_sync(0xffffffffu, n_nodes, 0);
        __syncwarp();
    }
    OCT_CLK(3);
    // drain (ORBextractor.cc:568-578): pop everything; popped entries pile up at the tail in reverse pop order
    const int total = heap_size;
    if (lane == 0) {
        int hs = heap_size;
        while (hs > 0) heap_pop(heap, hs);
    }
    __syncwarp();
    OCT_CLK(4);
    const int n_out = min(total, out_cap);
    for (int i = lane; i < n_out; i += 32) {
        const NodeRec nd = nodes[(uint32_t)heap[total - 1 - i]];
        const uint32_t* src = (nd.start >> 31) ? arena_b : arena_a;
        const int start = (int)(nd.start & 0x7fffffffu);
        uint32_t best = src[start];
        for (uint32_t j = 1; j < nd.cnt; ++j) {
            const uint32_t p = src[start + j];
            if (pt_r(p) > pt_r(best)) best = p;  // first maximum wins (strict >)
        }
        out[i] = best;
    }
    OCT_CLK(5);
    return total;
}

__global__ void __launch_bounds__(32 * OCT_WARPS_MAX) k_octree(const uint32_t* __restrict__ cell_pts, const int* __restrict__ cell_cnt,
                                                               uint32_t* __restrict__ arena_a, uint32_t* __restrict__ arena_b,
                                                               uint32_t* __restrict__ out_pts, int* __restrict__ out_cnt,
                                                               const __grid_constant__ Plan P, int n_images, int heap_cap, int pts_cap,
                                                               const int* __restrict__ only) {
    extern __shared__ __align__(16) unsigned long long oct_smem[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task = blockIdx.x * warps + warp;
    if (task >= n_images * P.n_levels) return;
    if (only && !only[task]) return;   // fallback launch: only the tasks k_octree_prep could not index with 16 bitsory
    // level-major task order so that the long level-0 tasks start first
    const int level = task / n_images, img = task - level * n_images;
    const LevelGeom& g = P.lv[level];
    // per-warp shared memory: heap (8 B) | nodes (16 B) | arena A | arena B (4 B each). heap_cap and pts_cap are even, so every
    // warp's block is 16-byte aligned; the heap proper starts at slot 1 (see heap_pop).
    const size_t per_warp = (size_t)heap_cap * 3 + (size_t)pts_cap;     // in 8-byte units
    unsigned long long* hbase = oct_smem + (size_t)warp * per_warp;
    unsigned long long* heap = hbase + 1;
    NodeRec* nodes = reinterpret_cast<NodeRec*>(hbase + heap_cap);
    uint32_t* sA = reinterpret_cast<uint32_t*>(hbase + (size_t)heap_cap * 3);
    uint32_t* sB = sA + pts_cap;

    const int* cnts = cell_cnt + (size_t)img * P.cells_per_image + g.cell_base;
    const uint32_t* cells = cell_pts + (size_t)img * P.cand_per_image + g.cand_off;
    const int n_cells = g.n_cols * g.n_rows;
    long long* clk = task == 0 ? g_oct_clk : nullptr;
    OCT_CLK(0);
    // total first: decides where the arenas live
    int M = 0;
    for (int c = lane; c < n_cells; c += 32) M += cnts[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M += __shfl_xor_sync(0xffffffffu, M, o);
    const bool in_smem = M <= pts_cap;
    uint32_t* A = in_smem ? sA : arena_a + (size_t)img * P.cand_per_image + g.cand_off;
    uint32_t* B = in_smem ? sB : arena_b + (size_t)img * P.cand_per_image + g.cand_off;

    // gather the per-cell lists into reference order (cell-row-major): 32 cells per round, one lane per cell
    int base = 0;
    for (int c0 = 0; c0 < n_cells; c0 += 32) {
        const int c = c0 + lane;
        const int k = c < n_cells ? cnts[c] : 0;
        int incl = k;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int dst0 = base + incl - k;
        const uint32_t* src = cells + (size_t)c * g.cell_cap;
        for (int j = 0; j < k; ++j) A[dst0 + j] = __ldg(src + j);
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    OCT_CLK(1);
    const int box_w = g.w - 2 * BORDER, box_h = g.h - 2 * BORDER;
    uint32_t* out = out_pts + (size_t)img * P.out_per_image + g.out_off;
    const int n = distribute_warp(A, B, M, box_w, box_h, g.n_ini, g.h_x, g.quota, g.out_cap, heap, nodes, out, lane, clk);
    if (lane == 0) out_cnt[(size_t)img * P.n_levels + level] = min(n, g.out_cap);
}

// ---------------------------------------------------------------------------------------------------------
// The sorted path (k_octree_prep + k_octree_replay), one CTA per (image, level) in each kernel: data-parallel phases (gather in
// reference order, path codes, counting sort by depth-T bucket, per-bucket sort), then one thread replays the heap on counts
// alone (octree_core.cuh) and its warp picks each surviving node's first-maximum-response point.
// ---------------------------------------------------------------------------------------------------------
#ifndef MCV_OC_THREADS
#define MCV_OC_THREADS 128
#endif
constexpr int OC_THREADS_BATCH = MCV_OC_THREADS;   // preparation kernel (data-parallel phases) when the tasks fill the GPU
constexpr int OC_THREADS_FEW = 512;                // ... and when there are only a few tasks (one triplet = 24): the task's own latency counts
constexpr int OR_THREADS = 32;               // replay kernel: one thread replays the heap, the warp picks the survivors' points

template <int OC_THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int* s_part, int& total) {   // OC_THREADS threads; s_part: OC_THREADS / 32 ints
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    __syncthreads();                      // s_part may still be read from a previous scan
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < OC_THREADS / 32; ++w) { const int t = s_part[w]; if (w < warp) base += t; tot += t; }
    total = tot;
    return base + incl - v;
}

// Sorts up to 8 (code, index) pairs held in registers (odd-even transposition, static indexing only).
__device__ __forceinline__ void sort8(uint32_t (&c)[8], uint16_t (&x)[8]) {
#pragma unroll
    for (int round = 0; round < 8; ++round) {
#pragma unroll
        for (int i = round & 1; i + 1 < 8; i += 2) {
            const bool sw = c[i + 1] < c[i];
            const uint32_t c0 = sw ? c[i + 1] : c[i], c1 = sw ? c[i] : c[i + 1];
            const uint16_t x0 = sw ? x[i + 1] : x[i], x1 = sw ? x[i] : x[i + 1];
            c[i] = c0; c[i + 1] = c1; x[i] = x0; x[i + 1] = x1;
        }
    }
}

// Two kernels per batch, one CTA per (image, level) task each:
//  k_octree_prep   (OC_THREADS threads): gather in reference order, path codes, counting sort by depth-T bucket, per-bucket
//                  sort; leaves `arena` (candidates in reference order), `scode` / `sidx` (path codes and original indices
//                  sorted by code) and the bucket prefix sums S (u16 x nb_pad per task) in global memory (L2-resident).
//  k_octree_replay (one warp): S into shared memory, thread 0 replays the heap on counts (octree_core.cuh), the warp picks every
//                  surviving node's first-maximum-response point. This kernel lasts as long as its longest heap replay, so it
//                  is kept as small as possible — 32 threads and ~7 KB of shared memory per task — and leaves the rest of the
//                  SM (registers above all: the one-kernel version pinned 64 threads x 48 registers per task for the whole
//                  replay, 98 % of the register file with all tasks resident) to the other stream's stencils.
//                  Measured on the level-0 task of a 640x480 image (434 nodes, B200, one triplet per call, scripts/r03/oct_clocks_b1.py):
//                  split loop 271 k cycles, drain 312 k, selection 18 k — a lone warp issues one dependent instruction every ~6.5
//                  cycles, and that, not memory, is the floor. Tried in round 2 and NOT kept (all bit-exact): a top-down pop that
//                  stops early instead of libstdc++'s sift-to-leaf + bubble-up (same arrays, but more instructions per level than the
//                  hand-scheduled loop in octree_core.cuh: drain 340 k); pushing sons that stay put without the sift loop (split 284 k);
//                  a warp-wide drain with pops following each other down the tree two levels apart, one lane each (drain 279 k for one
//                  task, but more instructions in total: the batched quadtree stage went 0.531 -> 0.549 ms); a pop by the whole warp
//                  (the hole's way through a 5-level subtree decided by 31 lanes at once and walked on ballot masks, moves one lane
//                  per level: split 279 k, drain 291 k — no better than the serial sift). Kept: the equal-key tail of the drain by
//                  lanes (below) and the child-count loads issued ahead of the pop, which matter when a level has fewer candidates
//                  than its quota — the reference's shipped single-level configuration, 1989 -> 1656 us per 512x512 triplet.
template <int OC_THREADS>
__global__ void __launch_bounds__(OC_THREADS) k_octree_prep(const uint32_t* __restrict__ cell_pts, const int* __restrict__ cell_cnt,
                                                            uint32_t* __restrict__ arena, uint32_t* __restrict__ scode_all,
                                                            uint16_t* __restrict__ sidx_all, uint16_t* __restrict__ S_all,
                                                            int* __restrict__ overflow,
                                                            const __grid_constant__ Plan P, int n_images, int nb_pad, int r0_words, int nb_smem) {
    extern __shared__ __align__(16) unsigned char oc_smem[];
    __shared__ int s_part[OC_THREADS / 32];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    const int task = blockIdx.x;                       // level-major: the long level-0 tasks start first
    const int level = task / n_images, img = task - level * n_images;
    const LevelGeom& lg = P.lv[level];
    uint32_t* cur = reinterpret_cast<uint32_t*>(oc_smem);      // histogram / scatter cursors (nb u32)
    uint16_t* S = reinterpret_cast<uint16_t*>(cur + r0_words);
    const size_t task_off = (size_t)img * P.cand_per_image + lg.cand_off;
    uint32_t* pts = arena + task_off;                  // candidates in reference order
    uint32_t* scode = scode_all + task_off;
    uint16_t* sidx = sidx_all + task_off;

    oct::Geom g;
    g.n_ini = lg.n_ini; g.h_x = lg.h_x; g.box_h = lg.h - 2 * BORDER; g.N = lg.quota; g.T = oct::tier_for(lg.n_ini);
    const int nb = lg.n_ini << (2 * g.T);
    const int bsh = 2 * (oct::DIGITS - g.T);
    long long* clk = task == 0 ? g_oct_clk : nullptr;
    if (clk && tid == 0) clk[0] = clock64();

    // -- gather: each thread owns a contiguous run of cells (reference order = cell-row-major, row-major inside a cell)
    const int* cnts = cell_cnt + (size_t)img * P.cells_per_image + lg.cell_base;
    const uint32_t* cells = cell_pts + task_off;
    const int n_cells = lg.n_cols * lg.n_rows;
    const int cpt = (n_cells + OC_THREADS - 1) / OC_THREADS;
    const int c_lo = min(tid * cpt, n_cells), c_hi = min(c_lo + cpt, n_cells);
    int mine = 0;
    for (int c = c_lo; c < c_hi; ++c) mine += cnts[c];
    for (int b = tid; b < nb; b += OC_THREADS) cur[b] = 0u;
    int M;
    int off = block_excl_scan<OC_THREADS>(mine, s_part, M);        // (its barriers also publish the zeroed histogram)
    if (M > 65535) {                                   // 16-bit indices: the legacy kernel takes this task
        if (tid == 0) overflow[task] = 1;
        return;
    }
    if (tid == 0) overflow[task] = 0;
    if (n_cells < nb_smem) {
        // balanced: the exclusive prefix of the cell counts goes to S (free until the histogram scan; M < 65536), then thread i
        // takes points i, i + OC_THREADS, ... and finds each one's cell by bisection. (A thread per run of cells leaves most
        // threads idle on the small levels: level 7 of 640x480 has 8 cells. ncu: barrier stall 7.4 warps per issue.)
        for (int c = c_lo; c < c_hi; ++c) { S[c] = (uint16_t)off; off += cnts[c]; }
        if (tid == 0) S[n_cells] = (uint16_t)M;
        __syncthreads();
        for (int i = tid; i < M; i += OC_THREADS) {
            int lo = 0, hi = n_cells;                  // last cell c with S[c] <= i (empty cells share their successor's prefix)
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if ((int)S[mid] <= i) lo = mid; else hi = mid; }
            const uint32_t p = __ldg(cells + (size_t)lo * lg.cell_cap + (i - (int)S[lo]));
            pts[i] = p;
            atomicAdd(&cur[min(oct::bucket_direct(p, g), nb - 1)], 1u);   // bucket histogram
        }
    } else {
        for (int c = c_lo; c < c_hi; ++c) {
            const int k = cnts[c];
            const uint32_t* src = cells + (size_t)c * lg.cell_cap;
            for (int j = 0; j < k; ++j) {
                const uint32_t p = __ldg(src + j);
                pts[off + j] = p;
                atomicAdd(&cur[min(oct::bucket_direct(p, g), nb - 1)], 1u);   // bucket histogram
            }
            off += k;
        }
    }
    __syncthreads();
    if (clk && tid == 0) clk[1] = clock64();
    // -- exclusive scan of the histogram: S (kept) and cur (scatter cursors)
    {
        const int bpt = (nb + OC_THREADS - 1) / OC_THREADS;
        const int b_lo = min(tid * bpt, nb), b_hi = min(b_lo + bpt, nb);
        int sum = 0;
        for (int b = b_lo; b < b_hi; ++b) sum += (int)cur[b];
        int tot;
        int run = block_excl_scan<OC_THREADS>(sum, s_part, tot);
        for (int b = b_lo; b < b_hi; ++b) { const int k = (int)cur[b]; S[b] = (uint16_t)run; cur[b] = (uint32_t)run; run += k; }
        if (tid == 0) S[nb] = (uint16_t)M;
    }
    __syncthreads();
    // -- scatter (code, index) into bucket order; arrival order inside a bucket is arbitrary, the per-bucket sort fixes it
#pragma unroll 4
    for (int i = tid; i < M; i += OC_THREADS) {
        const uint32_t pc = oct::path_code(__ldcg(pts + i), g);
        const uint32_t pos = atomicAdd(&cur[min((int)(pc >> bsh), nb - 1)], 1u);
        scode[pos] = pc;
        sidx[pos] = (uint16_t)i;
    }
    __syncthreads();
    // -- per-bucket sort by code (codes of distinct points are distinct): every tree node is now a contiguous range
    for (int b = tid; b < nb; b += OC_THREADS) {
        const uint32_t lo = S[b], n = S[b + 1] - lo;
        if (n < 2) continue;
        if (n <= 8) {
            uint32_t c[8]; uint16_t x[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { c[k] = k < (int)n ? __ldcg(scode + lo + k) : 0xffffffffu; x[k] = k < (int)n ? __ldcg(sidx + lo + k) : (uint16_t)0; }
            sort8(c, x);
#pragma unroll
            for (int k = 0; k < 8; ++k) if (k < (int)n) { scode[lo + k] = c[k]; sidx[lo + k] = x[k]; }
        } else {                                       // crowded bucket: plain insertion sort in global memory
            for (uint32_t i = lo + 1; i < lo + n; ++i) {
                const uint32_t c = __ldcg(scode + i);
                const uint16_t x = __ldcg(sidx + i);
                uint32_t j = i;
                while (j > lo && __ldcg(scode + j - 1) > c) { scode[j] = __ldcg(scode + j - 1); sidx[j] = __ldcg(sidx + j - 1); --j; }
                scode[j] = c; sidx[j] = x;
            }
        }
    }
    __syncthreads();
    if (clk && tid == 0) clk[2] = clock64();
    uint16_t* Sg = S_all + (size_t)task * nb_pad;
    for (int b = tid; b <= nb; b += OC_THREADS) Sg[b] = S[b];
}

#ifndef MCV_OR_MINB
#define MCV_OR_MINB 1
#endif
__global__ void __launch_bounds__(OR_THREADS, MCV_OR_MINB) k_octree_replay(const uint32_t* __restrict__ arena, const uint32_t* __restrict__ scode_all,
                                                              const uint16_t* __restrict__ sidx_all, const uint16_t* __restrict__ S_all,
                                                              uint32_t* __restrict__ out_pts, int* __restrict__ out_cnt,
                                                              const int* __restrict__ overflow, const __grid_constant__ Plan P, int n_images,
                                                              int nb_pad, int r0_words, int heap_alloc) {
    extern __shared__ __align__(16) unsigned char oc_smem[];
    __shared__ int s_total, s_tail;
    const int tid = threadIdx.x;
    const int task = blockIdx.x;                       // level-major: the long level-0 tasks start first
    if (overflow[task]) return;                        // the legacy kernel takes this task
    const int level = task / n_images, img = task - level * n_images;
    const LevelGeom& lg = P.lv[level];
    uint32_t* cur = reinterpret_cast<uint32_t*>(oc_smem);
    uint32_t* heap = cur + 1;                                  // heap - 1 is 16-byte aligned (oct::load2 / load4)
    uint32_t* nodes = cur + heap_alloc;
    uint16_t* S = reinterpret_cast<uint16_t*>(cur + r0_words);
    const size_t task_off = (size_t)img * P.cand_per_image + lg.cand_off;
    const uint32_t* pts = arena + task_off;            // candidates in reference order
    const uint32_t* scode = scode_all + task_off;
    const uint16_t* sidx = sidx_all + task_off;
    oct::Geom g;
    g.n_ini = lg.n_ini; g.h_x = lg.h_x; g.box_h = lg.h - 2 * BORDER; g.N = lg.quota; g.T = oct::tier_for(lg.n_ini);
    const int nb = lg.n_ini << (2 * g.T);
    long long* clk = task == 0 ? g_oct_clk : nullptr;
    const uint16_t* Sg = S_all + (size_t)task * nb_pad;
    if (clk && tid == 0) clk[6] = clock64();
    for (int b = tid; b <= nb; b += OR_THREADS) S[b] = Sg[b];
    __syncthreads();
    // -- serial heap replay on counts; the tail of the drain (all nodes down to one point: keys equal) by the warp, one lane per
    //    level of the heap's right-most spine (oct::equal_pop_plan)
    if (tid == 0) s_total = oct::replay(scode, S, g, heap, nodes, clk ? clk + 3 : nullptr, &s_tail);
    __syncthreads();
    const int total = s_total;
    {
        uint32_t* const hb = heap - 1;
        for (uint32_t sz = (uint32_t)s_tail; sz >= 1u; --sz) {
            const oct::EqPlan p = oct::equal_pop_plan(hb, sz, tid);
            __syncwarp();                                  // every lane has read the array as it was before this pop
            if (p.a0) hb[p.a0] = p.v0;
            if (p.a1) hb[p.a1] = p.v1;
            __syncwarp();
        }
    }
    if (clk && tid == 0) clk[4] = clock64();
    const int n_out = min(total, lg.out_cap);
    uint32_t* out = out_pts + (size_t)img * P.out_per_image + lg.out_off;
    for (int i = tid; i < n_out; i += OR_THREADS) out[i] = oct::select_best(heap[total - 1 - i], nodes, S, g.T, pts, sidx);
    if (tid == 0) {
        out_cnt[(size_t)img * P.n_levels + level] = n_out;
        if (clk) clk[5] = clock64();
    }
}

// Shared memory per warp: heap + nodes for `heap_cap` entries and two point arenas of pts_cap entries; as many warps per CTA
// (<= 4) as keep a CTA under ~56 KB so that several CTAs stay resident per SM.
static int oct_config(int heap_cap, int& pts_cap, int& warps, size_t& smem) {
    const size_t hn = (size_t)heap_cap * (sizeof(unsigned long long) + sizeof(NodeRec));
    if (hn > 200 * 1024) return -1;
    pts_cap = std::min(pts_cap, (int)((200 * 1024 - hn) / 8)) & ~1;
    const size_t per_warp = hn + (size_t)pts_cap * 8;
    warps = (int)std::min<size_t>(OCT_WARPS_MAX, std::max<size_t>(1, (56 * 1024) / per_warp));
    smem = per_warp * warps;
    return 0;
}

static int launch_octree_legacy(const Plan& P, const uint32_t* d_cell_pts, const int* d_cell_cnt, uint32_t* d_arena_a, uint32_t* d_arena_b,
                                uint32_t* d_out_pts, int* d_out_cnt, int n_images, const int* d_only, cudaStream_t s) {
    int heap_cap = 8, pts_cap = 64;
    for (int l = 0; l < P.n_levels; ++l) {
        heap_cap = std::max(heap_cap, (P.lv[l].out_cap + 4 + 1) & ~1);   // + slot 0 offset, even
        pts_cap = std::max(pts_cap, std::min(P.lv[l].cand_cap, 5 * P.lv[l].quota + 256));
    }
    int warps; size_t smem;
    if (oct_config(heap_cap, pts_cap, warps, smem)) return -1;
    // function attributes are per DEVICE and the call is cheap: set it on every launch that needs the opt-in (no process-wide cache)
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int tasks = n_images * P.n_levels;
    k_octree<<<(tasks + warps - 1) / warps, 32 * warps, smem, s>>>(d_cell_pts, d_cell_cnt, d_arena_a, d_arena_b, d_out_pts, d_out_cnt, P,
                                                                   n_images, heap_cap, pts_cap, d_only);
    return 1;
}

int launch_octree(const Plan& P, const uint32_t* d_cell_pts, const int* d_cell_cnt, uint32_t* d_arena_a, uint32_t* d_arena_b,
                  uint16_t* d_oct_idx, uint32_t* d_out_pts, int* d_out_cnt, int n_images, cudaStream_t s) {
#ifdef MCV_EXPERIMENTS
    // profiling aid, compiled only into experiment builds (MCV_NVCC_EXTRA=-DMCV_EXPERIMENTS): MCV_DEBUG_SKIP_OCTREE_AFTER=n stops
    // launching the quadtree after n calls — downstream then consumes the previous call's selection (valid for repeated input only)
    static const char* skip_env = getenv("MCV_DEBUG_SKIP_OCTREE_AFTER");
    static long skip_calls = 0;
    if (skip_env && ++skip_calls > atol(skip_env)) return 0;
#endif
    // shared-memory plan of k_octree_prep / k_octree_replay: R0 = cursors, later heap | nodes (r0_words u32) | S (nb_pad u16)
    int nb = 1, max_ini = 1, max_heap = 8;
    bool can_overflow = false;
    for (int l = 0; l < P.n_levels; ++l) {
        const LevelGeom& g = P.lv[l];
        max_ini = std::max(max_ini, g.n_ini);
        max_heap = std::max(max_heap, std::max(g.quota + 3, g.n_ini));
        nb = std::max(nb, g.n_ini << (2 * oct::tier_for(g.n_ini)));
        can_overflow |= g.cand_cap > 65535;
    }
    const int heap_alloc = (2 * (max_heap + 1) + 8 + 3) & ~3;              // oct::heap_pop's speculative reach (+ the slot before the heap)
    const int r0_words = (std::max(nb, heap_alloc + max_heap + 4) + 3) & ~3;
    const int nb_pad = (nb + 1 + 7) & ~7;
    const size_t smem = (size_t)r0_words * 4 + (size_t)nb_pad * 2;
    if (max_ini > oct::MAX_ROOTS || max_heap > 65000 || smem > 200 * 1024)   // panoramas / huge quotas: legacy kernel for everything
        return launch_octree_legacy(P, d_cell_pts, d_cell_cnt, d_arena_a, d_arena_b, d_out_pts, d_out_cnt, n_images, nullptr, s);
    // function attributes are per DEVICE: cache per device (16 is more than one box holds), set on first use there
    static bool configured[16] = {};
    static size_t configured_smem[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int di = dev & 15;
    if (!configured[di] || smem > configured_smem[di]) {
        if (smem > 48 * 1024) {
            cudaFuncSetAttribute(k_octree_prep<OC_THREADS_BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_octree_prep<OC_THREADS_FEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_octree_replay, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
        // many small CTAs that live as long as one thread's heap replay: ask for the largest shared-memory carve-out
        cudaFuncSetAttribute(k_octree_replay, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        configured[di] = true; configured_smem[di] = std::max(configured_smem[di], smem);   // benign race: the calls are idempotent
    }
    const int tasks = n_images * P.n_levels;
    int* d_overflow = d_out_cnt + tasks;
    // bucket prefix sums of every task: behind the sorted-index array (enqueue_extract reserves OCT_S_BYTES per task there)
    uint16_t* d_S = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(d_oct_idx) + (((size_t)P.cand_per_image * n_images * 2 + 255) & ~(size_t)255));
    if ((size_t)nb_pad * 2 > OCT_S_BYTES) return launch_octree_legacy(P, d_cell_pts, d_cell_cnt, d_arena_a, d_arena_b, d_out_pts, d_out_cnt, n_images, nullptr, s);
    if (tasks <= 2 * NUM_SMS)
        k_octree_prep<OC_THREADS_FEW><<<tasks, OC_THREADS_FEW, smem, s>>>(d_cell_pts, d_cell_cnt, d_arena_a, d_arena_b, d_oct_idx, d_S, d_overflow, P, n_images, OCT_S_BYTES / 2,
                                                                          r0_words, nb_pad);
    else
    k_octree_prep<OC_THREADS_BATCH><<<tasks, OC_THREADS_BATCH, smem, s>>>(d_cell_pts, d_cell_cnt, d_arena_a, d_arena_b, d_oct_idx, d_S, d_overflow, P, n_images, OCT_S_BYTES / 2, r0_words,
                                                  nb_pad);
    k_octree_replay<<<tasks, OR_THREADS, smem, s>>>(d_arena_a, d_arena_b, d_oct_idx, d_S, d_out_pts, d_out_cnt, d_overflow, P, n_images, OCT_S_BYTES / 2,
                                                    r0_words, heap_alloc);
    if (!can_overflow) return 2;
    return 2 + launch_octree_legacy(P, d_cell_pts, d_cell_cnt, d_arena_a, d_arena_b, d_out_pts, d_out_cnt, n_images, d_overflow, s);
}

int octree_debug_clocks(long long out[8]) { return cudaMemcpyFromSymbol(out, g_oct_clk, sizeof(long long) * 8) == cudaSuccess ? 0 : -1; }

// Static DistributeOctTree on caller points (mcv_orb_distribute_octree): one warp, points already in arena A.
__global__ void __launch_bounds__(32) k_octree_one(uint32_t* arena_a, uint32_t* arena_b, int M, int box_w, int box_h, int n_ini, float h_x,
                                                   int N, uint32_t* out, int* out_cnt, int out_cap, int heap_cap) {
    extern __shared__ __align__(16) unsigned long long oct_smem[];
    NodeRec* nodes = reinterpret_cast<NodeRec*>(oct_smem + heap_cap);
    const int n = distribute_warp(arena_a, arena_b, M, box_w, box_h, n_ini, h_x, N, out_cap, oct_smem + 1, nodes, out, threadIdx.x);
    if (threadIdx.x == 0) *out_cnt = min(n, out_cap);
}

int launch_octree_standalone(const uint32_t* d_pts, int n, int w_box, int h_box, int n_target, uint32_t* d_arena_a, uint32_t* d_arena_b,
                             uint32_t* d_out, int* d_out_cnt, int out_cap, cudaStream_t s) {
    // nIni / hX exactly as ORBextractor.cc:527-529
    const int n_ini = (int)roundf((float)w_box / (float)h_box);
    if (n_ini < 1) return -1;
    const float h_x = (float)w_box / (float)n_ini;
    const int heap_cap = (std::max(n_target + 4, n_ini + 4) + 4 + 1) & ~1;
    const size_t smem = (size_t)heap_cap * (sizeof(unsigned long long) + sizeof(NodeRec));
    if (smem > 220 * 1024) return -1;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_octree_one, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaMemcpyAsync(d_arena_a, d_pts, (size_t)n * 4, cudaMemcpyDeviceToDevice, s);
    k_octree_one<<<1, 32, smem, s>>>(d_arena_a, d_arena_b, n, w_box, h_box, n_ini, h_x, n_target, d_out, d_out_cnt, out_cap, heap_cap);
    return 1;
}

}  // namespace mcv
