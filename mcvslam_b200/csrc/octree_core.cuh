// Core of the quadtree quota distribution (ORBextractor::DistributeOctTree + ExtractorNode::DivideNode,
// ORBextractor.cc:469-580), written so that the SAME code runs in the sm_100a kernel (octree_kernels.cu) and, compiled by
// g++, in the CPU unit test (tests/cpp/octree_core_shim.cpp) — there is no GPU in the build container, so the order-defining
// logic is checked against the oracle on the host before it ever reaches the device.
//
// Idea. The reference splits the most populated node again and again (std::priority_queue keyed on the point count only) and
// each split re-partitions the node's points. But WHICH box a point falls into at every depth depends on the point and the
// level geometry alone, never on the split order. So every point gets its whole root-to-leaf path up front ("path code": root
// index, then one base-4 digit per depth, digit = (x >= mid_x) + 2 (y >= mid_y) — the order DivideNode pushes its sons in),
// the points are sorted by code once (data-parallel), and every node the reference can ever create is a contiguous range of
// that sorted array. What is left of the reference's loop is the heap itself: pop the top, look up the <= 4 child counts,
// push the non-empty ones. That replay is serial — which of several equally populated nodes is split next, and the order the
// survivors come out in, is decided by libstdc++'s __push_heap/__adjust_heap sift order, and keypoint order is API — but it
// now touches no point data: child counts of the shallow nodes come from a prefix-sum table over the depth-T buckets, those
// of deep nodes from a scan of a handful of sorted codes.
#pragma once
#include <stdint.h>
#ifndef __CUDACC__
#include <algorithm>
using std::max;
#endif

#ifdef __CUDACC__
#define OCT_HD __host__ __device__ __forceinline__
#else
#define OCT_HD inline
#endif

namespace mcv {
namespace oct {

constexpr int DIGITS = 14;        // path digits per point: 13 halvings separate any two distinct points with 12-bit coordinates
constexpr int MAX_ROOTS = 15;     // root index lives in bits 28..31 of the code
constexpr int MAX_TIER = 5;       // table depth: n_ini * 4^T buckets
constexpr int MAX_BUCKETS = 2048;

typedef unsigned long long u64;

struct Geom {
    int n_ini;    // roots (ORBextractor.cc:527)
    float h_x;    // root width (ORBextractor.cc:529)
    int box_h;    // maxY - minY
    int N;        // quota
    int T;        // table depth
};

OCT_HD int tier_for(int n_ini) {
    int T = MAX_TIER;
    while (T > 0 && (n_ini << (2 * T)) > MAX_BUCKETS) --T;
    return T;
}

OCT_HD int px(uint32_t p) { return (int)(p & 0xfffu); }
OCT_HD int py(uint32_t p) { return (int)((p >> 12) & 0xfffu); }
OCT_HD int pr(uint32_t p) { return (int)(p >> 24); }

OCT_HD float f_div(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
OCT_HD float f_mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;   // no contraction with a neighbouring add
    return r;
#endif
}

// Path of one point: root (ORBextractor.cc:545-548: init_nodes[kp.pt.x / hX]) and DIGITS DivideNode decisions
// (ORBextractor.cc:470-471 halves = ceil((float)extent / 2); :504-515 membership tests against n1.UR.x / n1.BR.y).
OCT_HD uint32_t path_code(uint32_t p, const Geom& g) {
    const int x = px(p), y = py(p);
    const int r = (int)(size_t)f_div((float)x, g.h_x);
    int ulx = (int)f_mul(g.h_x, (float)r), urx = (int)f_mul(g.h_x, (float)(r + 1)), uly = 0, bry = g.box_h;
    uint32_t code = (uint32_t)r;
#pragma unroll
    for (int k = 0; k < DIGITS; ++k) {
        const int mx = ulx + ((urx - ulx + 1) >> 1), my = uly + ((bry - uly + 1) >> 1);
        const bool right = x >= mx, bottom = y >= my;
        ulx = right ? mx : ulx; urx = right ? urx : mx;
        uly = bottom ? my : uly; bry = bottom ? bry : my;
        code = (code << 2) | (uint32_t)right | ((uint32_t)bottom << 1);
    }
    return code;
}
OCT_HD int bucket_of(uint32_t code, int T) { return (int)(code >> (2 * (DIGITS - T))); }
// bucket_of(path_code(p, g), T) without the digits below the table depth
OCT_HD int bucket_direct(uint32_t p, const Geom& g) {
    const int x = px(p), y = py(p);
    const int r = (int)(size_t)f_div((float)x, g.h_x);
    int ulx = (int)f_mul(g.h_x, (float)r), urx = (int)f_mul(g.h_x, (float)(r + 1)), uly = 0, bry = g.box_h;
    int b = r;
    for (int k = 0; k < g.T; ++k) {
        const int mx = ulx + ((urx - ulx + 1) >> 1), my = uly + ((bry - uly + 1) >> 1);
        const bool right = x >= mx, bottom = y >= my;
        ulx = right ? mx : ulx; urx = right ? urx : mx;
        uly = bottom ? my : uly; bry = bottom ? bry : my;
        b = (b << 2) | (int)right | ((int)bottom << 1);
    }
    return b;
}

// ---- heap entries: count << 16 | node id. The comparator looks at the count ONLY (ORBextractor.cc:530); the node id indexes
// `nodes`: depth << 28 | X, X = bucket prefix of the node (depth <= T: the node is buckets [X << 2(T-d), (X+1) << 2(T-d)))
//                                 or the start of its range in the sorted arrays (depth > T)
OCT_HD uint32_t e_cnt(uint32_t e) { return e >> 16; }
OCT_HD uint32_t e_id(uint32_t e) { return e & 0xffffu; }
OCT_HD bool e_less(uint32_t a, uint32_t b) { return (a | 0xffffu) < b; }   // a.count < b.count
OCT_HD int n_depth(uint32_t n) { return (int)(n >> 28); }
OCT_HD uint32_t n_x(uint32_t n) { return n & 0x0fffffffu; }
OCT_HD uint32_t make_node(int depth, uint32_t x) { return ((uint32_t)depth << 28) | x; }

// Heap array alignment contract: h - 1 is 16-byte aligned, so the child pair (odd, odd + 1) is one 8-byte access and the four
// grandchildren (4 hole + 3 ...) one 16-byte access. Speculative grandchildren may lie beyond the live heap: the array must
// hold 2 * max_size + 4 entries (their values are never used).
struct Pair2 { uint32_t x, y; };
struct Quad4 { uint32_t x, y, z, w; };
OCT_HD Pair2 load2(const uint32_t* h, int idx) {
#ifdef __CUDA_ARCH__
    const uint2 v = *reinterpret_cast<const uint2*>(h + idx);
    return Pair2{v.x, v.y};
#else
    return Pair2{h[idx], h[idx + 1]};
#endif
}
OCT_HD Quad4 load4(const uint32_t* h, int idx) {
#ifdef __CUDA_ARCH__
    const uint4 v = *reinterpret_cast<const uint4*>(h + idx);
    return Quad4{v.x, v.y, v.z, v.w};
#else
    return Quad4{h[idx], h[idx + 1], h[idx + 2], h[idx + 3]};
#endif
}

// std::push_heap after push_back (bits/stl_heap.h __push_heap): two ancestors are fetched per round trip
OCT_HD void heap_push(uint32_t* h, int& size, uint32_t value) {
    int hole = size++;
    while (hole > 0) {
        const int p1 = (hole - 1) >> 1;
        const int p2 = (max(p1, 1) - 1) >> 1;
        const uint32_t v1 = h[p1], v2 = h[p2];
        if (!e_less(v1, value)) break;
        h[hole] = v1;
        hole = p1;
        if (hole == 0 || !e_less(v2, value)) break;
        h[hole] = v2;
        hole = p2;
    }
    h[hole] = value;
}

// std::pop_heap + pop_back (bits/stl_heap.h __pop_heap -> __adjust_heap -> __push_heap). The popped element is left at
// h[size - 1] like std::pop_heap leaves it, and returned. (Fetching grandchildren speculatively, two levels per round trip, was
// measured on B200 and bought nothing: the single thread is bound by instructions per level, not by the load latency.)
OCT_HD uint32_t heap_pop(uint32_t* h, int& size) {
    const uint32_t top = h[0];
    if (size > 1) {
        const int len = size - 1;
        const uint32_t value = h[len];
        h[len] = top;
        const int lim = (len - 1) >> 1;   // __adjust_heap: nodes below lim have two children
        // The heap replay runs on ONE thread, where every instruction costs its full dependent-issue latency, so the loop is
        // written for the shortest dependent chain per level. With hb[i + 1] = h[i] and A = address of the hole's slot, the
        // hole's children are the 8-byte pair at 2A - hb, and the next hole is that pair's left or right word.
        int hole;
#ifdef __CUDA_ARCH__
        {
            const uint32_t base = (uint32_t)__cvta_generic_to_shared(h - 1);
            uint32_t A = base + 4u;                          // hb[1] = h[0]
            const uint32_t A_end = base + 4u * (uint32_t)lim;    // hole < lim  <=>  A <= A_end
            asm volatile(
                "{\n\t"
                ".reg .pred p, q;\n\t"
                ".reg .u32 pa, pb, l, r, t, v, nb;\n\t"
                "neg.s32 nb, %2;\n\t"
                "setp.gt.u32 q, %0, %1;\n\t"
                "@q bra OCT_SIFT_DONE_%=;\n"
                "OCT_SIFT_LOOP_%=:\n\t"
                "mad.lo.u32 pa, %0, 2, nb;\n\t"            // left child slot
                "add.u32 pb, pa, 4;\n\t"
                "ld.shared.v2.u32 {l, r}, [pa];\n\t"
                "or.b32 t, r, 0xffff;\n\t"
                "setp.lt.u32 p, t, l;\n\t"                 // comp(right, left): take the left child
                "selp.u32 v, l, r, p;\n\t"
                "st.shared.u32 [%0], v;\n\t"
                "selp.u32 %0, pa, pb, p;\n\t"
                "setp.le.u32 q, %0, %1;\n\t"
                "@q bra OCT_SIFT_LOOP_%=;\n"
                "OCT_SIFT_DONE_%=:\n\t"
                "}"
                : "+r"(A) : "r"(A_end), "r"(base) : "memory");
            hole = (int)((A - base) >> 2) - 1;
        }
#else
        {
            uint32_t* const hb = h - 1;
            int H = 1;
            while (H <= lim) {
                const Pair2 p = load2(hb, 2 * H);
                const bool r = !e_less(p.y, p.x);         // comp(right, left) false: take the right child
                hb[H] = r ? p.y : p.x;
                H = 2 * H + (int)r;
            }
            hole = H - 1;                                 // == "secondChild" of __adjust_heap after every step
        }
#endif
        if ((len & 1) == 0 && hole == ((len - 2) >> 1)) {
            const int c = 2 * (hole + 1);
            h[hole] = h[c - 1];
            hole = c - 1;
        }
        while (hole > 0) {
            const int parent = (hole - 1) >> 1;
            const uint32_t pv = h[parent];
            if (!e_less(pv, value)) break;
            h[hole] = pv;
            hole = parent;
        }
        h[hole] = value;
    }
    --size;
    return top;
}

// ---- drain of a heap whose elements all compare equal (every node holds one point: the tail of every drain, and the whole
// drain when a level has fewer candidates than its quota — the reference's shipped single-level configuration) ----
// With equal keys std::pop_heap decides nothing from the data: comp(right, left) is false, so the hole runs down the right-most
// spine 1, 3, 7, ... (1-based slots 2^(j+1) - 1) while a right child exists, takes a lone left child if that is what is left,
// and the displaced last element stays where the hole ends (no parent is smaller). One pop is therefore the same data movement
// for every heap of that size, and the spine's levels can move at the same time: "lane" j plans the write(s) of spine level j
// from the array as it is BEFORE the pop; all plans are read first, then written. hb = h - 1, s = current size; the popped
// element is left in slot s like heap_pop leaves it.
struct EqPlan { uint32_t a0, v0, a1, v1; };              // up to two writes hb[a] = v (a == 0: none)
constexpr int EQ_LANES = 17;                             // spine levels of a heap of < 2^16 elements, + the slot-s writer
OCT_HD EqPlan equal_pop_plan(const uint32_t* hb, uint32_t s, int lane) {
    EqPlan p{0u, 0u, 0u, 0u};
    if (lane >= EQ_LANES) return p;
    const uint32_t len = s - 1u;                          // elements that stay
    const uint32_t pj = (2u << lane) - 1u, pn = (4u << lane) - 1u, pp = (1u << lane) - 1u;   // this, next and previous spine slot
    if (pn <= len) { p.a0 = pj; p.v0 = hb[pn]; }          // the right child moves up
    else if (pj <= len) {                                 // the spine ends here
        if (2u * pj == len) { p.a0 = pj; p.v0 = hb[len]; p.a1 = len; p.v1 = hb[s]; }   // lone left child moves up, the value takes its slot
        else { p.a0 = pj; p.v0 = hb[s]; }
    } else if (lane == 0 || pp <= len) { p.a0 = s; p.v0 = hb[1]; }   // first lane past the spine: the popped top goes to slot s
    return p;
}

// start of a node's range in the sorted arrays
template <typename ST>
OCT_HD int node_lo(uint32_t node, const ST* S, int T) {
    const int d = n_depth(node);
    return d <= T ? (int)S[n_x(node) << (2 * (T - d))] : (int)n_x(node);
}

// The serial part (ORBextractor.cc:549-578): roots, split loop, drain. `scode` = path codes sorted ascending (only deep nodes
// read it), `S` = exclusive prefix sums of the bucket histogram (n_ini << 2T entries + 1). Returns the number of final nodes; the i-th popped node
// (= i-th output keypoint) is left at heap[total - 1 - i] (count << 16 | id, node description in nodes[id]) — once the caller has
// drained the equal-key tail, if it asked for one.
template <typename ST>
OCT_HD int replay(const uint32_t* scode, const ST* S, const Geom& g, uint32_t* heap, uint32_t* nodes, long long* clk_split_done = nullptr,
                  int* equal_tail = nullptr) {
    const int T = g.T;
    int size = 0, n_nodes = 0;
    for (int r = 0; r < g.n_ini; ++r) {   // ORBextractor.cc:549-555: empty roots are dropped
        const uint32_t c = (uint32_t)S[(r + 1) << (2 * T)] - (uint32_t)S[r << (2 * T)];
        if (c) { nodes[n_nodes] = make_node(0, (uint32_t)r); heap_push(heap, size, (c << 16) | (uint32_t)n_nodes); ++n_nodes; }
    }
    if (size == 0) return 0;
    // every iteration grows the heap or descends one depth, and depths are bounded by DIGITS: the guard is only reached on
    // input the reference itself would spin on forever (coincident points)
    for (int guard = 16 * g.N + 256; size < g.N && guard > 0; --guard) {
        const uint32_t top = heap[0];
        const uint32_t cnt = e_cnt(top), id = e_id(top);
        if (cnt == 1) break;
        const uint32_t node = nodes[id];
        const int d = n_depth(node);
        if (d >= DIGITS) break;
        // The child counts depend on the popped node alone, so their loads are ISSUED before the pop and consumed after it: the
        // pop's dependent sift (hundreds of cycles on a lone thread) then covers the latency of the sorted codes, which live in
        // global memory (an L2 round trip per deep split: 749 of the 1084 splits of a 512x512 single-level image).
        uint32_t c0, c1, c2, c3, x0, xs;   // child counts; child k's X = x0 + k * xs (table) or running range start (deep)
        if (d < T) {
            const int sh = 2 * (T - 1 - d);
            x0 = n_x(node) << 2;
            const ST* s = S + ((size_t)x0 << sh);
            const uint32_t s0 = s[0], s1 = s[(size_t)1 << sh], s2 = s[(size_t)2 << sh], s3 = s[(size_t)3 << sh], s4 = s[(size_t)4 << sh];
            heap_pop(heap, size);
            c0 = s1 - s0; c1 = s2 - s1; c2 = s3 - s2; c3 = s4 - s3;
            xs = 1;
        } else {
            x0 = d == T ? (uint32_t)S[n_x(node)] : n_x(node);
            const int sh = 2 * (DIGITS - 1 - d);
            constexpr uint32_t PRE = 6;   // codes fetched ahead (deep nodes hold few points)
            uint32_t pre[PRE];
#pragma unroll
            for (uint32_t j = 0; j < PRE; ++j) pre[j] = j < cnt ? scode[x0 + j] : 0u;
            heap_pop(heap, size);
            unsigned long long acc = 0;   // four 16-bit counters
#pragma unroll
            for (uint32_t j = 0; j < PRE; ++j) if (j < cnt) acc += 1ull << (16 * ((pre[j] >> sh) & 3u));
            for (uint32_t j = PRE; j < cnt; ++j) acc += 1ull << (16 * ((scode[x0 + j] >> sh) & 3u));
            c0 = (uint32_t)acc & 0xffffu; c1 = (uint32_t)(acc >> 16) & 0xffffu; c2 = (uint32_t)(acc >> 32) & 0xffffu; c3 = (uint32_t)(acc >> 48);
            xs = 0;
        }
        // sons in DivideNode's order n1..n4 (ORBextractor.cc:516-519); the popped node's slot is reused by the first one
        bool reuse = true;
        uint32_t at = x0;
#define OCT_PUSH_SON(ck, k)                                                     \
        if (ck) {                                                               \
            const uint32_t nid = reuse ? id : (uint32_t)n_nodes++;              \
            reuse = false;                                                      \
            nodes[nid] = make_node(d + 1, xs ? x0 + k : at);                    \
            heap_push(heap, size, ((ck) << 16) | nid);                          \
        }                                                                       \
        at += ck;
        OCT_PUSH_SON(c0, 0) OCT_PUSH_SON(c1, 1) OCT_PUSH_SON(c2, 2) OCT_PUSH_SON(c3, 3)
#undef OCT_PUSH_SON
    }
    const int total = size;
#ifdef __CUDA_ARCH__
    if (clk_split_done) *clk_split_done = clock64();
#endif
    // with `equal_tail` the serial drain stops as soon as the top holds one point — then every remaining element does — and leaves
    // that many elements in heap[0, *equal_tail) for the caller's equal-key drain (equal_pop_plan)
    if (equal_tail) {
        while (size > 0 && e_cnt(heap[0]) > 1) heap_pop(heap, size);
        *equal_tail = size;
        return total;
    }
    while (size > 0) heap_pop(heap, size);
    return total;
}

// First maximum response in the node's vKeys order (ORBextractor.cc:571-577: strict >, and vKeys keeps the order of
// vToDistributeKeys through every stable partition) = max of (response, -original index) over the node's range.
template <typename ST>
OCT_HD uint32_t select_best(uint32_t entry, const uint32_t* nodes, const ST* S, int T, const uint32_t* pts, const uint16_t* sidx) {
    const int lo = node_lo(nodes[e_id(entry)], S, T), cnt = (int)e_cnt(entry);
    uint32_t best_key = 0, best = 0;
    for (int j = 0; j < cnt; ++j) {
        const uint32_t i = sidx[lo + j];
        const uint32_t p = pts[i];
        const uint32_t key = ((uint32_t)pr(p) << 16) | (0xffffu - i);
        if (j == 0 || key > best_key) { best_key = key; best = p; }
    }
    return best;
}

}  // namespace oct
}  // namespace mcv
