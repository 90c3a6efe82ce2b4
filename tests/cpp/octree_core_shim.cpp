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

// The kernel's warp-wide drain (drain_pipelined in octree_kernels.cu) with the 32 lanes emulated: same start rule, same per-lane
// oct::pop_step, lanes visited in the given order inside a step (the hardware gives none) — `order` 0: ascending, 1: descending,
// 2: a different shuffle every step. Also checks the property the pipeline rests on: inside one step no lane reads or writes a
// slot that another lane writes. Returns the number of violations (0 expected); popped[k] = k-th popped entry.
static int drain_emulated(uint32_t* heap, int n, uint32_t* popped, int order) {
    uint32_t* const hb = heap - 1;
    uint32_t hole[32] = {0}, len[32] = {0}, value[32] = {0};
    bool active[32] = {false};
    int next = 0, since = 2, violations = 0;
    uint32_t rng = 12345u;
    while (true) {
        if (next < n && since >= 2) {
            const uint32_t size = (uint32_t)(n - next);
            bool hazard = false;
            for (int l = 0; l < 32; ++l) hazard |= active[l] && is_ancestor_or_self(hole[l], size);
            if (!hazard) {
                const int l = next & 31;
                if (active[l]) ++violations;
                popped[next] = hb[1];
                if (size > 1) { value[l] = hb[size]; len[l] = size - 1u; hole[l] = 1u; active[l] = true; }
                ++next; since = 0;
            }
        }
        bool any = false;
        for (int l = 0; l < 32; ++l) any |= active[l];
        if (!any) { if (next >= n) break; since = 2; continue; }
        // slots touched in this step: a lane in flight reads 2 * hole and 2 * hole + 1 and writes `hole`
        for (int a = 0; a < 32; ++a) for (int b = 0; b < 32; ++b) {
            if (a == b || !active[a] || !active[b]) continue;
            const uint32_t w = hole[a];
            if (w == hole[b] || ((w == 2 * hole[b] || w == 2 * hole[b] + 1) && 2 * hole[b] <= len[b])) ++violations;
        }
        int perm[32];
        for (int l = 0; l < 32; ++l) perm[l] = order == 1 ? 31 - l : l;
        if (order == 2) for (int l = 31; l > 0; --l) { rng = rng * 1664525u + 1013904223u; std::swap(perm[l], perm[(rng >> 8) % (uint32_t)(l + 1)]); }
        for (int i = 0; i < 32; ++i) { const int l = perm[i]; if (active[l]) active[l] = pop_step(hb, hole[l], len[l], value[l]); }
        ++since;
    }
    return violations;
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
    if (drain_order < 0) {
        const int total = replay(scode.data(), S16.data(), g, heap, nodes.data());
        for (int i = 0; i < total && i < out_cap; ++i) out[i] = select_best(heap[total - 1 - i], nodes.data(), S16.data(), g.T, pts, sidx.data());
        return total;
    }
    const int total = replay_split(scode.data(), S16.data(), g, heap, nodes.data());
    std::vector<uint32_t> popped(total + 1);
    if (drain_emulated(heap, total, popped.data(), drain_order) != 0) return -3;
    for (int i = 0; i < total && i < out_cap; ++i) out[i] = select_best(popped[i], nodes.data(), S16.data(), g.T, pts, sidx.data());
    return total;
}

extern "C" int octcore_distribute(const uint32_t* pts, int M, int box_w, int box_h, int N, uint32_t* out, int out_cap) {
    return distribute_impl(pts, M, box_w, box_h, N, out, out_cap, -1);
}
// split loop + the kernel's pipelined drain, lanes emulated (drain_order: see drain_emulated)
extern "C" int octcore_distribute_pipelined(const uint32_t* pts, int M, int box_w, int box_h, int N, uint32_t* out, int out_cap, int drain_order) {
    return distribute_impl(pts, M, box_w, box_h, N, out, out_cap, drain_order);
}

// Heap primitives alone against libstdc++: a random heap of n entries with counts in [1, max_count] (small ranges = tie-heavy) is
// built with std::push_heap / oct::heap_push, then drained three ways — std::pop_heap, oct::heap_pop, the emulated pipeline.
// Returns the number of disagreements (entry order or array contents).
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
    std::vector<uint32_t> pipe_store(mine_store), popped(n + 1);
    uint32_t* pipe = pipe_store.data() + 1;
    bad += drain_emulated(pipe, n, popped.data(), (int)(seed % 3u));
    for (int k = 0; k < n; ++k) {
        std::pop_heap(ref.begin(), ref.end(), cmp);
        const uint32_t a = ref.back(); ref.pop_back();
        const uint32_t b = heap_pop(mine, size);
        bad += a != b; bad += a != popped[k];
        for (size_t i = 0; i < ref.size(); ++i) bad += ref[i] != mine[i];
    }
    return bad;
}
