// Brute-force Hamming 2-NN on the 5th-generation tensor cores (sm_100a: tcgen05.mma kind::i8, accumulators in TMEM, operands
// staged by TMA into 128-byte-swizzled shared memory). Replaces the O(Q*T) loops of Matcher::KnnMatch / cv::BFMatcher
// (src/Matcher.cpp:134-144,245-308) for every problem large enough to fill a tile.
//
// Why this is the b1 path of the north star: on sm_100a there is no binary tensor-core instruction — ptxas lowers
// `mma.sync.m16n8k256.b1.and.popc` to eight IMMA.16832.U8.U8 on bit planes it expands with MOVM.U4TO8 on every call
// (cuobjdump of scripts/exp/b1_mma_probe.cu, profiles/r02_b1_verdict.md), i.e. AND-popc IS an int8 dot product of 0/1 vectors.
// Here the expansion is done ONCE per descriptor (k_expand_pm1: bit -> +-1 int8, 256 bytes per row) and the distance table is
// an int8 GEMM:    <a', b'> = (#equal bits) - (#different bits) = 256 - 2 Hamming(a, b)      (exact in the s32 accumulator).
//
// k_knn2_tc: one CTA owns 128 queries (UMMA M) and walks its train range in tiles of 256 rows (UMMA N); K = 256 bytes = two
// swizzle atoms = 8 instructions of K = 32 per tile. Warp 0 = TMA producer (3-stage ring of 64 KB train tiles, the query tile
// is loaded once), warp 1 = MMA issuer and TMEM owner (two 128 x 256 s32 accumulators = all 512 columns, so the epilogue of
// tile i overlaps the MMAs of tile i + 1), warps 2..17 = epilogue: thread = one query row (TMEM lane), four warps per lane
// quarter take 64 columns each (two tcgen05.ld.32x32b.x32 in flight together; the TMEM buffer is handed back as soon as they
// land, before the scan). The 2-NN needs no key per pair: a thread sees its columns in ascending train index, so a column can
// only enter the top-2 when its dot product is STRICTLY larger than the current second best's; one max-reduction per 32
// columns decides that, and only then the (distance << 22 | index) keys are formed (about 2 ln T times per row). Results are
// the same lexicographic (distance, index) top-2 as k_knn2_bf.
#include <cuda.h>

#include "engine.h"
#include "tma.cuh"

namespace mcv {

constexpr int TC_M = 128;
constexpr int TC_N = 256;
constexpr int TC_KB = 128;                 // bytes of K per shared-memory tile row (one swizzle atom)
constexpr int TC_STAGES = 3;                // 3 x 64 KB train tiles + 32 KB of queries: a TMA load has two tiles of MMA time to land
constexpr int TC_EPI_WARPS = 16;            // four warps per TMEM lane quarter, 64 columns of every tile each
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_EPI_COLS = TC_N / (TC_EPI_WARPS / 4);
constexpr unsigned TC_A_TILE = TC_M * TC_KB;            // 16 KB per K half
constexpr unsigned TC_B_TILE = TC_N * TC_KB;            // 32 KB per K half
constexpr unsigned TC_STAGE_BYTES = 2 * TC_B_TILE;
constexpr unsigned TC_SMEM_BYTES = 2 * TC_A_TILE + TC_STAGES * TC_STAGE_BYTES + 1024 /*alignment*/ + 256 /*barriers*/;
static_assert(TC_SMEM_BYTES <= 232448, "shared memory per CTA");
#ifndef MCV_TC_SCAN_GROUP
#define MCV_TC_SCAN_GROUP 8
#endif
constexpr int TC_SCAN_GROUP = MCV_TC_SCAN_GROUP;   // columns per group of the epilogue's slow path
constexpr int TC_IDX_BITS = 22;            // same key as k_knn2_bf: distance << 22 | trainIdx
constexpr unsigned TC_SENT = 0xffffffffu;

__device__ __forceinline__ unsigned tc_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_arrive(unsigned b) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"(b) : "memory");
}
// K-major operand tile with 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO); descriptor version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(unsigned smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor, kind::i8: D = s32 (2 << 4), A and B signed 8 bit (1 << 7, 1 << 10), both K-major, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t TC_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void umma_i8(unsigned tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, int (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
                 "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
                   "=r"(v[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tc_top2(unsigned& k0, unsigned& k1, unsigned key) {
    k1 = min(k1, max(k0, key));
    k0 = min(k0, key);
}

// descriptors (32 bytes) -> 256 int8 of +-1 (bit set -> -1), bit b of byte j at position 8 j + b. One thread per 32-bit word.
__global__ void __launch_bounds__(256) k_expand_pm1(const uint8_t* __restrict__ a, long long words_a, int8_t* __restrict__ out_a,
                                                    const uint8_t* __restrict__ b, long long words_b, int8_t* __restrict__ out_b) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const uint8_t* src = a; int8_t* out = out_a;
    if (i >= words_a) { i -= words_a; if (i >= words_b) return; src = b; out = out_b; }
    const unsigned w = reinterpret_cast<const unsigned*>(src)[i];
    uint4 o[2];
    unsigned* ow = reinterpret_cast<unsigned*>(o);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const unsigned nib = (w >> (4 * k)) & 0xfu;
        const unsigned bits = (nib * 0x00204081u) & 0x01010101u;    // bit j of the nibble -> byte j
        ow[k] = 0x01010101u ^ (bits * 0xfeu);                          // 0 -> 0x01 (+1), 1 -> 0xff (-1)
    }
    uint4* dst = reinterpret_cast<uint4*>(out + i * 32);
    dst[0] = o[0]; dst[1] = o[1];
}

struct TcArgs {
    const int32_t* q_img;      // per pair: image (slab) of the query map / train map; nullptr = slab 0
    const int32_t* t_img;
    const int32_t* q_cnt;      // per image row counts; nullptr = nq / nt for every pair
    const int32_t* t_cnt;
    int nq, nt;
    int per_split, n_splits;
    int q_stride;              // query rows per pair in the outputs
    int train_offset;
    unsigned* part;            // [pair][split][q_stride][2] keys (n_splits > 1)
    int32_t* idx;              // [pair][q_stride][2]
    int32_t* dist;
};

// grid = (ceil(max nq / 128), n_splits, n_pairs)
__global__ void __launch_bounds__(TC_THREADS, 1) k_knn2_tc(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t,
                                                           const TcArgs A) {
    extern __shared__ uint8_t smem_raw[];
    const int pair = blockIdx.z;
    const int qi = A.q_img ? A.q_img[pair] : 0, ti = A.t_img ? A.t_img[pair] : 0;
    const int nq = A.q_cnt ? A.q_cnt[qi] : A.nq, nt = A.t_cnt ? A.t_cnt[ti] : A.nt;
    const int q0 = blockIdx.x * TC_M;
    if (q0 >= nq) return;                                            // block-uniform, before any barrier or TMEM allocation
    const unsigned base = (tc_smem_u32(smem_raw) + 1023u) & ~1023u;
    const unsigned sA = base, sB = base + 2 * TC_A_TILE, sBar = sB + TC_STAGES * TC_STAGE_BYTES;
    // barriers: a_full | full[S] | empty[S] | tmem_full[2] | tmem_empty[2]; then the TMEM base address
    const unsigned bar_a = sBar, bar_full = sBar + 8, bar_empty = bar_full + 8 * TC_STAGES, bar_tfull = bar_empty + 8 * TC_STAGES, bar_tempty = bar_tfull + 16;
    const unsigned tmem_slot = bar_tempty + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t_begin = blockIdx.y * A.per_split, t_end = min(nt, t_begin + A.per_split);
    const int n_tiles = t_end > t_begin ? (t_end - t_begin + TC_N - 1) / TC_N : 0;

    if (threadIdx.x == 0) {
        mbar_init(bar_a, 1);
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // all 512 TMEM columns: two 128 x 256 s32 accumulators (one CTA per SM: 161 KB of shared memory)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0 && n_tiles > 0) {
            mbar_expect_tx(bar_a, 2 * TC_A_TILE);
            tma_load_3d(sA, &map_q, 0, q0, qi, bar_a);
            tma_load_3d(sA + TC_A_TILE, &map_q, TC_KB, q0, qi, bar_a);
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i % TC_STAGES;
                if (i >= TC_STAGES) mbar_wait(bar_empty + 8 * s, ((i / TC_STAGES) - 1) & 1);
                mbar_expect_tx(bar_full + 8 * s, TC_STAGE_BYTES);
                const int row = t_begin + i * TC_N;
                tma_load_3d(sB + s * TC_STAGE_BYTES, &map_t, 0, row, ti, bar_full + 8 * s);
                tma_load_3d(sB + s * TC_STAGE_BYTES + TC_B_TILE, &map_t, TC_KB, row, ti, bar_full + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && n_tiles > 0) {
            mbar_wait(bar_a, 0);
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i % TC_STAGES, b = i & 1;
                if (i >= 2) mbar_wait(bar_tempty + 8 * b, ((i >> 1) - 1) & 1);
                mbar_wait(bar_full + 8 * s, (i / TC_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned d_tmem = tmem_base + (unsigned)(b * TC_N);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t da = umma_desc_sw128(sA + kb * TC_A_TILE), db = umma_desc_sw128(sB + s * TC_STAGE_BYTES + kb * TC_B_TILE);
#pragma unroll
                    for (int k = 0; k < 4; ++k)    // 32 bytes of K per instruction: the start address advances inside the swizzle atom
                        umma_i8(d_tmem, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), (kb | k) ? 1u : 0u);
                }
                umma_commit(bar_empty + 8 * s);     // the stage is free once these MMAs have read it
                umma_commit(bar_tfull + 8 * b);     // the accumulator is complete
            }
        }
    } else {
        // epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 = query rows; one row per thread, TC_EPI_COLS columns of every tile
        const int quarter = warp & 3, part = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        unsigned k0 = TC_SENT, k1 = TC_SENT;
        int thr = -100000;                          // dot product of the current second best: only a strictly larger one can enter
        // One 32-column chunk of this row. Fast path: four group maxima, one compare. Slow path (some column beats the current
        // second best): only the groups of eight whose maximum qualifies are walked — a warp takes the slow path when ANY of its
        // 32 rows qualifies, which on short train walks (5000 rows) is nearly every chunk, so its cost is what matters there
        // (all 32 columns unconditionally: 5000 x 5000 = 55 us; by groups: see profiles/r02_matching_pipes.md).
        auto scan = [&](const int (&v)[32], int idx0) {
            constexpr int GS = TC_SCAN_GROUP, NG = 32 / GS;
            int g[NG];
#pragma unroll
            for (int i = 0; i < NG; ++i) {
                g[i] = v[GS * i];
#pragma unroll
                for (int j = 1; j < GS; ++j) g[i] = max(g[i], v[GS * i + j]);
            }
            int m = g[0];
#pragma unroll
            for (int i = 1; i < NG; ++i) m = max(m, g[i]);
            if (m > thr) {
#pragma unroll
                for (int i = 0; i < NG; ++i) {
                    if (g[i] > thr) {
#pragma unroll
                        for (int j = GS * i; j < GS * i + GS; ++j) {
                            if (v[j] > thr && idx0 + j < t_end) {
                                tc_top2(k0, k1, ((unsigned)((256 - v[j]) >> 1) << TC_IDX_BITS) | (unsigned)(idx0 + j));
                                if (k1 != TC_SENT) thr = 256 - 2 * (int)(k1 >> TC_IDX_BITS);
                            }
                        }
                    }
                }
            }
        };
        static_assert(TC_EPI_COLS == 64, "the epilogue keeps its two 32-column loads of a tile in flight together");
        for (int i = 0; i < n_tiles; ++i) {
            const int b = i & 1;
            mbar_wait(bar_tfull + 8 * b, (i >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int col0 = part * TC_EPI_COLS;
            const int idx_base = t_begin + i * TC_N + col0;
            const unsigned taddr = tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)(b * TC_N + col0);
            int va[32], vb[32];
            tmem_ld32(taddr, va);
            tmem_ld32(taddr + 32u, vb);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            // the accumulator is in registers: hand the TMEM buffer back before scanning
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(bar_tempty + 8 * b);
            scan(va, idx_base);
            scan(vb, idx_base + 32);
        }
        // [parts - 1][128][2] keys of the other column parts, in the first train stage: every MMA has completed (the last tile's
        // accumulator was awaited above), so nothing reads the stages any more
        unsigned* sk = reinterpret_cast<unsigned*>(smem_raw + (sB - tc_smem_u32(smem_raw)));
        if (part > 0) { sk[((part - 1) * TC_M + row) * 2] = k0; sk[((part - 1) * TC_M + row) * 2 + 1] = k1; }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory");                               // epilogue warps only
        const int q = q0 + row;
        if (part == 0 && q < nq) {
#pragma unroll
            for (int p = 0; p < TC_EPI_WARPS / 4 - 1; ++p) { tc_top2(k0, k1, sk[(p * TC_M + row) * 2]); tc_top2(k0, k1, sk[(p * TC_M + row) * 2 + 1]); }
            if (A.n_splits > 1) {
                unsigned* o = A.part + (((size_t)pair * A.n_splits + blockIdx.y) * A.q_stride + q) * 2;
                o[0] = k0; o[1] = k1;
            } else {
                const unsigned mask = (1u << TC_IDX_BITS) - 1;
                const size_t o = ((size_t)pair * A.q_stride + q) * 2;
                A.idx[o] = k0 == TC_SENT ? -1 : (int)(k0 & mask) + A.train_offset;
                A.dist[o] = k0 == TC_SENT ? 0x7fffffff : (int)(k0 >> TC_IDX_BITS);
                A.idx[o + 1] = k1 == TC_SENT ? -1 : (int)(k1 & mask) + A.train_offset;
                A.dist[o + 1] = k1 == TC_SENT ? 0x7fffffff : (int)(k1 >> TC_IDX_BITS);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// merge of the per-split partial keys: grid covers n_pairs * q_stride rows
__global__ void __launch_bounds__(256) k_knn2_tc_merge(const TcArgs A, int n_pairs) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (long long)n_pairs * A.q_stride) return;
    const int pair = (int)(r / A.q_stride), q = (int)(r % A.q_stride);
    const int qi = A.q_img ? A.q_img[pair] : 0;
    if (q >= (A.q_cnt ? A.q_cnt[qi] : A.nq)) return;
    unsigned k0 = TC_SENT, k1 = TC_SENT;
    for (int s = 0; s < A.n_splits; ++s) {
        const unsigned* p = A.part + (((size_t)pair * A.n_splits + s) * A.q_stride + q) * 2;
        tc_top2(k0, k1, p[0]); tc_top2(k0, k1, p[1]);
    }
    const unsigned mask = (1u << TC_IDX_BITS) - 1;
    const size_t o = ((size_t)pair * A.q_stride + q) * 2;
    A.idx[o] = k0 == TC_SENT ? -1 : (int)(k0 & mask) + A.train_offset;
    A.dist[o] = k0 == TC_SENT ? 0x7fffffff : (int)(k0 >> TC_IDX_BITS);
    A.idx[o + 1] = k1 == TC_SENT ? -1 : (int)(k1 & mask) + A.train_offset;
    A.dist[o + 1] = k1 == TC_SENT ? 0x7fffffff : (int)(k1 >> TC_IDX_BITS);
}

// (256 bytes, rows, slabs) int8 view of an expanded descriptor array, box = 128 bytes x box_rows x 1, 128-byte swizzle
static bool tc_encode_map(CUtensorMap* m, const int8_t* base, int rows, int slabs, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dims[3] = {256, (cuuint64_t)rows, (cuuint64_t)slabs}, strides[2] = {256, (cuuint64_t)rows * 256};
    const cuuint32_t box[3] = {(cuuint32_t)TC_KB, (cuuint32_t)box_rows, 1u}, estr[3] = {1u, 1u, 1u};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<int8_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static void tc_splits(int nq, int nt, int n_pairs, int& n_splits, int& per_split) {
    const long long q_tiles = (long long)((nq + TC_M - 1) / TC_M) * n_pairs;
    const int t_tiles = std::max(1, (nt + TC_N - 1) / TC_N);
    // one CTA walks the whole train range unless that leaves most SMs idle on a long walk: then the range is split (and merged)
    n_splits = 1;
    if (q_tiles < NUM_SMS && t_tiles > 12) n_splits = (int)std::min<long long>(t_tiles / 4, (2 * NUM_SMS + q_tiles - 1) / q_tiles);
    n_splits = std::max(1, n_splits);
    per_split = ((t_tiles + n_splits - 1) / n_splits) * TC_N;
    n_splits = std::max(1, (nt + per_split - 1) / per_split);
}

// single problems: from 2^23 pairs up (B200: 2000 x 2000 = 23 us on the integer pipe vs 31 us here — the expansion launch and the
// pipeline fill dominate; 5000 x 5000 = 68 us vs 42 us)
bool knn2_tc_usable(int nq, int nt) { return nq >= 1 && nt >= 2 && (long long)nq * nt >= (1 << 23) && nt <= (1 << TC_IDX_BITS); }

// scratch of one call: expanded queries | expanded train rows | partial keys
size_t knn2_tc_scratch_bytes(int nq_rows, int nt_rows, int nq, int nt, int n_pairs) {
    int n_splits, per_split;
    tc_splits(nq, nt, n_pairs, n_splits, per_split);
    const size_t part = n_splits > 1 ? (size_t)n_pairs * n_splits * nq * 8 : 0;
    return (((size_t)nq_rows * 256 + 1023) & ~(size_t)1023) + (((size_t)nt_rows * 256 + 1023) & ~(size_t)1023) + part + 1024;
}

// Single problem (d_t_same == false) or a batch of pairs over one descriptor array [n_images][cap][32] with per-image counts.
//   single: d_q [nq][32], d_t [nt][32]                       -> idx / dist [nq][2]
//   pairs:  d_q == d_t == descriptors of n_images x cap rows -> idx / dist [n_pairs][cap][2], pair p = (q_img[p], t_img[p])
int launch_knn2_tc(const uint8_t* d_q, int nq, const uint8_t* d_t, int nt, int train_offset, int32_t* d_idx, int32_t* d_dist, void* d_scratch,
                   int n_images, const int32_t* d_counts, const int32_t* d_pair_q, const int32_t* d_pair_t, int n_pairs, cudaStream_t s) {
    const bool pairs = n_images > 0;
    if (!pairs) n_pairs = 1;
    if (n_pairs <= 0 || nq <= 0) return 0;
    const int q_rows = pairs ? n_images * nq : nq, t_rows = pairs ? n_images * nt : nt;     // pairs: nq == nt == cap
    int8_t* e_q = reinterpret_cast<int8_t*>(d_scratch);
    int8_t* e_t = pairs ? e_q : e_q + (((size_t)q_rows * 256 + 1023) & ~(size_t)1023);
    unsigned* d_part = reinterpret_cast<unsigned*>((pairs ? e_q + (((size_t)q_rows * 256 + 1023) & ~(size_t)1023) : e_t + (((size_t)t_rows * 256 + 1023) & ~(size_t)1023)));
    static bool configured[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured[dev & 15]) {
        if (cudaFuncSetAttribute(k_knn2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES) != cudaSuccess) return -1;
        configured[dev & 15] = true;
    }
    const long long wq = (long long)q_rows * 8, wt = pairs ? 0 : (long long)t_rows * 8;
    k_expand_pm1<<<(unsigned)((wq + wt + 255) / 256), 256, 0, s>>>(d_q, wq, e_q, d_t, wt, e_t);
    CUtensorMap mq, mt;
    if (!tc_encode_map(&mq, e_q, nq, pairs ? n_images : 1, TC_M) || !tc_encode_map(&mt, e_t, nt, pairs ? n_images : 1, TC_N)) return -1;
    TcArgs A{};
    A.q_img = d_pair_q; A.t_img = d_pair_t; A.q_cnt = pairs ? d_counts : nullptr; A.t_cnt = pairs ? d_counts : nullptr;
    A.nq = nq; A.nt = nt; A.q_stride = nq; A.train_offset = train_offset;
    tc_splits(nq, nt, n_pairs, A.n_splits, A.per_split);
    A.part = d_part; A.idx = d_idx; A.dist = d_dist;
    k_knn2_tc<<<dim3((nq + TC_M - 1) / TC_M, A.n_splits, n_pairs), TC_THREADS, TC_SMEM_BYTES, s>>>(mq, mt, A);
    if (A.n_splits == 1) return 2;
    k_knn2_tc_merge<<<(unsigned)(((long long)n_pairs * nq + 255) / 256), 256, 0, s>>>(A, n_pairs);
    return 3;
}

}  // namespace mcv
