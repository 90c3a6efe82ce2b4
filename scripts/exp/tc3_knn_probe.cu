// Probe: tensor-core Hamming 2-NN with the A operand (the CTA's 128 queries, expanded to +-1 int8 on the fly) in TMEM
// (tcgen05.mma ... [a_tmem], b_desc: "TS" form) and N = 192 tiles: the MMA then reads only B from shared memory (64 B/clk) next to
// TMA's 64 B/clk of writes = 128 B/clk, the shared-memory bandwidth, instead of 160 B/clk with A in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tc3_knn_probe tc3_knn_probe.cu && ./tc3_knn_probe --time
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int TC_M = 128;
constexpr int TC_N = 192;            // train rows per tile: 2 x 192 accumulator columns + 64 columns of A = 448 <= 512
constexpr int TC_KB = 128;
#ifndef TC_STAGES_N
#define TC_STAGES_N 4
#endif
constexpr int TC_STAGES = TC_STAGES_N;
constexpr int EPI_WARPS = 12;        // three per TMEM lane quarter, 64 columns of every tile each
constexpr int TC_THREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_COLS = 64;
constexpr unsigned B_TILE_BYTES = TC_N * TC_KB;           // 24 KB per K half
constexpr unsigned STAGE_BYTES = 2 * B_TILE_BYTES;        // 48 KB
constexpr unsigned SMEM_BYTES = TC_STAGES * STAGE_BYTES + 1024 + 256;
constexpr unsigned A_COL0 = 2 * TC_N;                     // TMEM column of the A operand
constexpr int IDX_BITS = 22;
constexpr unsigned SENT = 0xffffffffu;
#ifndef A_BYTE_ORDER_BE
#define A_BYTE_ORDER_BE 0
#endif

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned b, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned b, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.release.cta.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned b) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.release.cta.shared::cta.b64 st, [%0];\n\t}" ::"r"(b) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned b, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\nW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\nD_%=:\n\t}" ::"r"(b), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(unsigned smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffff) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
// A from TMEM, B from shared memory
__device__ __forceinline__ void umma_i8_ts(unsigned tmem_d, unsigned tmem_a, uint64_t db, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
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
__device__ __forceinline__ void tmem_st32(unsigned taddr, const unsigned (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
                 "%23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
                   "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
                   "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
                   "r"(v[31]) : "memory");
}
__host__ __device__ __forceinline__ void top2_update(unsigned& k0, unsigned& k1, unsigned key) { const unsigned hi = k0 > key ? k0 : key; k1 = k1 < hi ? k1 : hi; k0 = k0 < key ? k0 : key; }

__device__ __forceinline__ unsigned expand_nibble(unsigned nib) {      // 4 bits -> 4 bytes of +-1 (bit set -> -1), bit j -> byte j
    const unsigned bits = (nib * 0x00204081u) & 0x01010101u;
    unsigned w = 0x01010101u ^ (bits * 0xfeu);
#if A_BYTE_ORDER_BE
    w = __byte_perm(w, 0, 0x0123);
#endif
    return w;
}

__global__ void k_expand_pm1(const uint8_t* __restrict__ d, int n, int8_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 8) return;
    const unsigned w = reinterpret_cast<const unsigned*>(d)[i];
    uint4 o[2];
    unsigned* ow = reinterpret_cast<unsigned*>(o);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const unsigned nib = (w >> (4 * k)) & 0xfu;
        const unsigned bits = (nib * 0x00204081u) & 0x01010101u;
        ow[k] = 0x01010101u ^ (bits * 0xfeu);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)i * 32);
    dst[0] = o[0]; dst[1] = o[1];
}

// grid = (ceil(nq / 128), n_splits); q = the PACKED query descriptors (32 bytes per row), map_t over the expanded train rows
__global__ void __launch_bounds__(TC_THREADS, 1) k_knn2_tc(const uint8_t* __restrict__ q, const __grid_constant__ CUtensorMap map_t, int nq, int nt,
                                                           int per_split, unsigned* __restrict__ part) {
    extern __shared__ uint8_t smem_raw[];
    const unsigned base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const unsigned sB = base, sBar = sB + TC_STAGES * STAGE_BYTES;
    const unsigned bar_a = sBar, bar_full = sBar + 8, bar_empty = bar_full + 8 * TC_STAGES, bar_tfull = bar_empty + 8 * TC_STAGES, bar_tempty = bar_tfull + 16;
    const unsigned tmem_slot = bar_tempty + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * TC_M;
    const int t_begin = blockIdx.y * per_split, t_end = min(nt, t_begin + per_split);
    const int n_tiles = (t_end - t_begin + TC_N - 1) / TC_N;

    if (threadIdx.x == 0) {
        mbar_init(bar_a, 8);                                   // the eight warps that write A into TMEM
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i % TC_STAGES;
                if (i >= TC_STAGES) mbar_wait(bar_empty + 8 * s, ((i / TC_STAGES) - 1) & 1);
                mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
                const int row = t_begin + i * TC_N;
                tma_load_2d(sB + s * STAGE_BYTES, &map_t, 0, row, bar_full + 8 * s);
                tma_load_2d(sB + s * STAGE_BYTES + B_TILE_BYTES, &map_t, TC_KB, row, bar_full + 8 * s);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(bar_a, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int i = 0; i < n_tiles; ++i) {
                const int s = i % TC_STAGES, b = i & 1;
                if (i >= 2) mbar_wait(bar_tempty + 8 * b, ((i >> 1) - 1) & 1);
                mbar_wait(bar_full + 8 * s, (i / TC_STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned d_tmem = tmem_base + (unsigned)(b * TC_N);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t db = umma_desc_sw128(sB + s * STAGE_BYTES + kb * B_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)    // K = 32 bytes per instruction = 8 TMEM columns of A
                        umma_i8_ts(d_tmem, tmem_base + A_COL0 + (unsigned)((kb * 4 + k) * 8), db + (uint64_t)(k * 2), (kb | k) ? 1u : 0u);
                }
                umma_commit(bar_empty + 8 * s);
                umma_commit(bar_tfull + 8 * b);
            }
        }
    } else {
        const int quarter = warp & 3, part_i = (warp - 2) >> 2;          // 0..2
        const int row = quarter * 32 + lane;
        // A operand: parts 0 and 1 expand half of their row's descriptor (16 bytes -> 128 int8 -> 32 TMEM columns) each
        if (part_i < 2) {
            unsigned a[32];
            uint4 d = make_uint4(0u, 0u, 0u, 0u);
            if (q0 + row < nq) d = __ldg(reinterpret_cast<const uint4*>(q + (size_t)(q0 + row) * 32) + part_i);
            const unsigned dw[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int w = 0; w < 4; ++w)
#pragma unroll
                for (int k = 0; k < 8; ++k) a[w * 8 + k] = q0 + row < nq ? expand_nibble((dw[w] >> (4 * k)) & 0xfu) : 0u;
            tmem_st32(tmem_base + ((unsigned)(quarter * 32) << 16) + A_COL0 + (unsigned)(part_i * 32), a);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_a);
        }
        unsigned k0 = SENT, k1 = SENT;
        int thr = -100000;
        auto scan = [&](const int (&v)[32], int idx0) {
            int g[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                g[i] = v[8 * i];
#pragma unroll
                for (int j = 1; j < 8; ++j) g[i] = max(g[i], v[8 * i + j]);
            }
            if (max(max(g[0], g[1]), max(g[2], g[3])) > thr) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (g[i] > thr) {
#pragma unroll
                        for (int j = 8 * i; j < 8 * i + 8; ++j) {
                            if (v[j] > thr && idx0 + j < t_end) {
                                top2_update(k0, k1, ((unsigned)((256 - v[j]) >> 1) << IDX_BITS) | (unsigned)(idx0 + j));
                                if (k1 != SENT) thr = 256 - 2 * (int)(k1 >> IDX_BITS);
                            }
                        }
                    }
                }
            }
        };
        for (int i = 0; i < n_tiles; ++i) {
            const int b = i & 1;
            mbar_wait(bar_tfull + 8 * b, (i >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int col0 = part_i * EPI_COLS;
            const int idx_base = t_begin + i * TC_N + col0;
            const unsigned taddr = tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)(b * TC_N + col0);
            int va[32], vb[32];
            tmem_ld32(taddr, va);
            tmem_ld32(taddr + 32u, vb);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
            scan(va, idx_base);
            scan(vb, idx_base + 32);
        }
        unsigned* sk = reinterpret_cast<unsigned*>(smem_raw + (sB - smem_u32(smem_raw)));
        if (part_i > 0) { sk[((part_i - 1) * TC_M + row) * 2] = k0; sk[((part_i - 1) * TC_M + row) * 2 + 1] = k1; }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
        if (part_i == 0 && q0 + row < nq) {
#pragma unroll
            for (int p = 0; p < EPI_WARPS / 4 - 1; ++p) { top2_update(k0, k1, sk[(p * TC_M + row) * 2]); top2_update(k0, k1, sk[(p * TC_M + row) * 2 + 1]); }
            unsigned* o = part + ((size_t)blockIdx.y * nq + q0 + row) * 2;
            o[0] = k0; o[1] = k1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) { __syncwarp(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory"); }
}

__global__ void k_merge(const unsigned* part, int nq, int n_splits, int* idx, int* dist) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    unsigned k0 = SENT, k1 = SENT;
    for (int s = 0; s < n_splits; ++s) { const unsigned* p = part + ((size_t)s * nq + q) * 2; top2_update(k0, k1, p[0]); top2_update(k0, k1, p[1]); }
    const unsigned mask = (1u << IDX_BITS) - 1;
    idx[2 * q] = k0 == SENT ? -1 : (int)(k0 & mask); dist[2 * q] = k0 == SENT ? 0x7fffffff : (int)(k0 >> IDX_BITS);
    idx[2 * q + 1] = k1 == SENT ? -1 : (int)(k1 & mask); dist[2 * q + 1] = k1 == SENT ? 0x7fffffff : (int)(k1 >> IDX_BITS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_map(CUtensorMap* m, const int8_t* base, int rows, int box_rows) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr; cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[2] = {256, (cuuint64_t)rows}, strides[1] = {256};
    const cuuint32_t box[2] = {(cuuint32_t)TC_KB, (cuuint32_t)box_rows}, estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Runner {
    int8_t *eq = nullptr, *et = nullptr; unsigned* part = nullptr; int *idx = nullptr, *dist = nullptr;
    uint8_t *dq = nullptr, *dt = nullptr;
    int nq, nt, n_splits, per_split;
    void setup(const std::vector<uint8_t>& q, const std::vector<uint8_t>& t, int nq_, int nt_, int splits_hint) {
        nq = nq_; nt = nt_;
        CK(cudaMalloc(&dq, (size_t)nq * 32)); CK(cudaMalloc(&dt, (size_t)nt * 32));
        CK(cudaMemcpy(dq, q.data(), (size_t)nq * 32, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt, t.data(), (size_t)nt * 32, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&eq, (size_t)nq * 256)); CK(cudaMalloc(&et, (size_t)nt * 256));
        const int q_tiles = (nq + TC_M - 1) / TC_M, t_tiles = (nt + TC_N - 1) / TC_N;
        n_splits = splits_hint > 0 ? splits_hint : std::max(1, std::min((148 + q_tiles - 1) / q_tiles, t_tiles));
        per_split = ((t_tiles + n_splits - 1) / n_splits) * TC_N;
        n_splits = (nt + per_split - 1) / per_split;
        CK(cudaMalloc(&part, (size_t)n_splits * nq * 8)); CK(cudaMalloc(&idx, (size_t)nq * 8)); CK(cudaMalloc(&dist, (size_t)nq * 8));
        CK(cudaFuncSetAttribute(k_knn2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    }
    void run(bool expand_t) {
        if (expand_t) k_expand_pm1<<<(nt * 8 + 255) / 256, 256>>>(dt, nt, et);
        CUtensorMap mq, mt;
        if (!make_map(&mt, et, nt, TC_N)) { printf("tensor map encode failed\n"); exit(1); }
        k_knn2_tc<<<dim3((nq + TC_M - 1) / TC_M, n_splits), TC_THREADS, SMEM_BYTES>>>(dq, mt, nq, nt, per_split, part);
        k_merge<<<(nq + 255) / 256, 256>>>(part, nq, n_splits, idx, dist);
    }
    void release() { cudaFree(dq); cudaFree(dt); cudaFree(eq); cudaFree(et); cudaFree(part); cudaFree(idx); cudaFree(dist); }
};

static int check(int nq, int nt, bool low, int splits_hint, unsigned seed) {
    std::vector<uint8_t> q((size_t)nq * 32), t((size_t)nt * 32);
    srand(seed);
    for (auto& b : q) b = (uint8_t)(low ? rand() & 3 : rand());
    for (auto& b : t) b = (uint8_t)(low ? rand() & 3 : rand());
    Runner r; r.setup(q, t, nq, nt, splits_hint);
    r.run(true);
    CK(cudaDeviceSynchronize());
    std::vector<int> idx((size_t)nq * 2), dist((size_t)nq * 2);
    CK(cudaMemcpy(idx.data(), r.idx, idx.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(dist.data(), r.dist, dist.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int i = 0; i < nq; ++i) {
        unsigned k0 = SENT, k1 = SENT;
        for (int j = 0; j < nt; ++j) {
            int d = 0;
            for (int w = 0; w < 4; ++w) d += __builtin_popcountll(((const uint64_t*)&q[(size_t)i * 32])[w] ^ ((const uint64_t*)&t[(size_t)j * 32])[w]);
            top2_update(k0, k1, ((unsigned)d << IDX_BITS) | (unsigned)j);
        }
        const int e[4] = {(int)(k0 & ((1u << IDX_BITS) - 1)), (int)(k0 >> IDX_BITS), k1 == SENT ? -1 : (int)(k1 & ((1u << IDX_BITS) - 1)), k1 == SENT ? 0x7fffffff : (int)(k1 >> IDX_BITS)};
        if (idx[2 * i] != e[0] || dist[2 * i] != e[1] || idx[2 * i + 1] != e[2] || dist[2 * i + 1] != e[3]) {
            if (bad < 5) printf("  q %d: got (%d,%d) (%d,%d) want (%d,%d) (%d,%d)\n", i, idx[2 * i], dist[2 * i], idx[2 * i + 1], dist[2 * i + 1], e[0], e[1], e[2], e[3]);
            ++bad;
        }
    }
    printf("check nq %d nt %d low %d splits %d (per_split %d): %d mismatches\n", nq, nt, (int)low, r.n_splits, r.per_split, bad);
    r.release();
    return bad;
}

static void timeit(int nq, int nt, int reps) {
    std::vector<uint8_t> q((size_t)nq * 32), t((size_t)nt * 32);
    srand(5);
    for (auto& b : q) b = (uint8_t)rand();
    for (auto& b : t) b = (uint8_t)rand();
    Runner r; r.setup(q, t, nq, nt, 0);
    r.run(true); CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int pass = 0; pass < 2; ++pass) {
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) r.run(pass == 0);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
        printf("time nq %d nt %d (%s): %.3f ms, %.3e pairs/s, %.1f TOP/s int8\n", nq, nt, pass == 0 ? "incl. train expansion" : "train pre-expanded", ms,
               (double)nq * nt / (ms * 1e-3), (double)nq * nt * 512 / (ms * 1e-3) / 1e12);
    }
    r.release();
}

int main(int argc, char** argv) {
    int bad = 0;
    bad += check(128, 256, false, 1, 1);
    bad += check(300, 1000, false, 0, 2);
    bad += check(257, 4100, true, 0, 3);
    bad += check(2000, 2003, false, 0, 4);
    bad += check(1, 2, false, 0, 5);
    bad += check(1000, 5000, true, 1, 6);
    printf("total mismatching checks: %d\n", bad);
    if (argc > 1 && !strcmp(argv[1], "--time")) {
        timeit(2000, 2000, 50);
        timeit(5000, 5000, 50);
        timeit(131072, 1 << 20, 2);
    }
    return bad != 0;
}
