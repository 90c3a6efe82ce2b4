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
                                                               const __grid_constant__ Plan P, int n_images, int heap_cap, int pts_cap) {
    extern __shared__ __align__(16) unsigned long long oct_smem[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int task = blockIdx.x * warps + warp;
    if (task >= n_images * P.n_levels) return;
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

int launch_octree(const Plan& P, const uint32_t* d_cell_pts, const int* d_cell_cnt, uint32_t* d_arena_a, uint32_t* d_arena_b,
                  uint32_t* d_out_pts, int* d_out_cnt, int n_images, cudaStream_t s) {
    int heap_cap = 8, pts_cap = 64;
    for (int l = 0; l < P.n_levels; ++l) {
        heap_cap = std::max(heap_cap, (P.lv[l].out_cap + 4 + 1) & ~1);   // + slot 0 offset, even
        pts_cap = std::max(pts_cap, std::min(P.lv[l].cand_cap, 5 * P.lv[l].quota + 256));
    }
    int warps; size_t smem;
    if (oct_config(heap_cap, pts_cap, warps, smem)) return -1;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    const int tasks = n_images * P.n_levels;
    k_octree<<<(tasks + warps - 1) / warps, 32 * warps, smem, s>>>(d_cell_pts, d_cell_cnt, d_arena_a, d_arena_b, d_out_pts, d_out_cnt, P,
                                                                   n_images, heap_cap, pts_cap);
    return 1;
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
