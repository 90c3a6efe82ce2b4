// Test infrastructure: runs the quadtree core that the sm_100a kernel uses (mcvslam_b200/csrc/octree_core.cuh) on the CPU,
// with the kernel's data-parallel phases replayed sequentially (and the bucket scatter deliberately in REVERSE arrival order,
// since the kernel's atomics give no order), so tests/test_octree_core.py can check it against the oracle without a GPU.
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../mcvslam_b200/csrc/octree_core.cuh"

using namespace mcv::oct;
using std::max;

static void equal_drain_emulated(uint32_t* heap, int size, int order) {
    uint32_t* const hb = heap - 1;
    for (uint32_t s = (uint32_t)size; s >= 1u; --s) {
        EqPlan plan[32];
        for (int l = 0; l < 32; ++l) plan[l] = equal_pop_plan(hb, s, l);
        for (int i = 0; i < 32; ++i) {
            const EqPlan& p = plan[order ? 31 - i : i];
            if (p.a0) hb[p.a0] = p.v0;
            if (p.a1) hb[p.a1] = p.v1;
        }
    }
}

static int distribute_impl(const uint32_t* pts, int M, int box_w, int box_h, int N, uint32_t* out, int out_cap, int drain_order) {
    const int n_ini = (int)roundf((float)box_w / (float)box_h);   // ORBextractor.cc:527
    if (n_ini < 1 || n_ini > MAX_ROOTS || M > 65535) return -1;
    Geom g{n_ini, (float)box_w / (float)n_ini, box_h, N, tier_for(n_ini)};
    const int nb = n_ini << (2 * g.T);
    std::vector<uint32_t> code(M), S(nb + 1, 0), cur(nb, 0), scode(M);
    std::vector<uint16_t> sidx(M);
    for (int i = 0; i < M; ++i) { code[i] = path_code(pts[i], g); if (bucket_of(code[i], g.T) >= nb || bucket_direct(pts[i], g) != bucket_of(code[i], g.T)) return -2; ++cur[bucket_of(code[i], g.T)]; }
    for (int b = 0; b < nb; ++b) { S[b + 1] = S[b] + cur[b]; cur[b] = S[b]; }
    for (int i = M - 1; i >= 0; --i) { const uint32_t pos = cur[bucket_of(code[i], g.T)]++; scode[pos] = code[i]; sidx[pos] = (uint16_t)i; }
    for (int b = 0; b < nb; ++b)   // per-bucket insertion sort, as the kernel does
        for (uint32_t i = S[b] + 1; i < S[b + 1]; ++i) {
            const uint32_t c = scode[i]; const uint16_t x = sidx[i];
            uint32_t j = i;
            while (j > S[b] && scode[j - 1] > c) { scode[j] = scode[j - 1]; sidx[j] = sidx[j - 1]; --j; }
            scode[j] = c; sidx[j] = x;
        }
    std::vector<uint32_t> heap_store(2 * (std::max(N + 3, n_ini) + 1) + 8 + 2), nodes(std::max(N + 3, n_ini) + 4);
    uint32_t* heap = heap_store.data() + 1;
    std::vector<uint16_t> S16(S.begin(), S.end());   // the kernel keeps the table in 16 bits
    // drain_order >= 0: the kernel's form — serial replay up to the equal-key tail, then equal_pop_plan with the lanes emulated
    // (all plans made from the array as it is before the pop, then written; written in ascending or descending lane order)
    int total;
    if (drain_order < 0) total = replay(scode.data(), S16.data(), g, heap, nodes.data());
    else {
        int tail = 0;
        total = replay(scode.data(), S16.data(), g, heap, nodes.data(), nullptr, &tail);
        equal_drain_emulated(heap, tail, drain_order);
    }
    for (int i = 0; i < total && i < out_cap; ++i) out[i] = select_best(heap[total - 1 - i], nodes.data(), S16.data(), g.T, pts, sidx.data());
    return total;
}

extern "C" int octcore_distribute(const uint32_t* pts, int M, int box_w, int box_h, int N, uint32_t* out, int out_cap) {
    return distribute_impl(pts, M, box_w, box_h, N, out, out_cap, -1);
}

// Heap primitives alone against libstdc++: a random heap of n entries with counts in [1, max_count] (small ranges = tie-heavy) is
// built with std::push_heap / oct::heap_push, then drained by std::pop_heap and by oct::heap_pop (the replay's hand-scheduled sift).
// Returns the number of disagreements (entry order or array contents).
// split loop + serial drain down to the equal-key tail + the kernel's equal-key drain with its lanes emulated
extern "C" int octcore_distribute_kernel_form(const uint32_t* pts, int M, int box_w, int box_h, int N, uint32_t* out, int out_cap, int drain_order) {
    return distribute_impl(pts, M, box_w, box_h, N, out, out_cap, drain_order);
}

extern "C" int octcore_heap_check(unsigned seed, int n, int max_count) {
    auto cmp = [](uint32_t a, uint32_t b) { return (a >> 16) < (b >> 16); };
    std::vector<uint32_t> ref, mine_store(2 * n + 16, 0u);
    uint32_t* mine = mine_store.data() + 1;
    int size = 0, bad = 0;
    uint32_t rng = seed * 2654435761u + 1u;
    for (int i = 0; i < n; ++i) {
        rng = rng * 1664525u + 1013904223u;
        const uint32_t e = ((1u + (rng >> 10) % (uint32_t)max_count) << 16) | (uint32_t)i;
        ref.push_back(e); std::push_heap(ref.begin(), ref.end(), cmp);
        heap_push(mine, size, e);
    }
    for (int i = 0; i < n; ++i) bad += ref[i] != mine[i];
    if (max_count == 1) {                                  // equal keys: the lane-parallel drain must leave exactly what n x heap_pop leaves
        std::vector<uint32_t> eq_store(mine_store), ser_store(mine_store);
        equal_drain_emulated(eq_store.data() + 1, n, (int)(seed & 1u));
        int sz = n;
        while (sz > 0) heap_pop(ser_store.data() + 1, sz);
        for (int i = 0; i < n; ++i) bad += eq_store[1 + i] != ser_store[1 + i];
    }
    for (int k = 0; k < n; ++k) {
        std::pop_heap(ref.begin(), ref.end(), cmp);
        const uint32_t a = ref.back(); ref.pop_back();
        const uint32_t b = heap_pop(mine, size);
        bad += a != b;
        for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != mine[i];
    }
    return bad;
}
