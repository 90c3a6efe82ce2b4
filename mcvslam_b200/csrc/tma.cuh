// TMA (cp.async.bulk.tensor) + mbarrier helpers shared by the kernels that stage image windows in shared memory, and the
// host-side tensor-map encoder for the per-level (pitch, h, n_images) u8 views of a pyramid-shaped buffer.
// Rules learned on B200 (scripts/exp/tma_probe.cu): the start coordinate of the byte dimension must be a multiple of 16 (an
// unaligned one raises "illegal instruction"), the inner box extent a multiple of 16 bytes, the shared-memory destination
// 128-byte aligned; everything outside the tensor is zero-filled and still counted in the transaction bytes.
#pragma once
#include <cuda.h>
#include "engine.h"

#ifndef MCV_TMA_L2PROMO
#define MCV_TMA_L2PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_128B
#endif

namespace mcv {

__device__ __forceinline__ void tma_load_3d(unsigned dst_smem, const CUtensorMap* map, int c0, int c1, int c2, unsigned mbar_smem) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst_smem), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(mbar_smem) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned mbar_smem, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_smem), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar_smem, unsigned bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.release.cta.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(mbar_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar_smem, unsigned parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n"
        "MCV_MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MCV_MBAR_DONE_%=;\n\t"
        "bra MCV_MBAR_WAIT_%=;\n"
        "MCV_MBAR_DONE_%=:\n\t"
        "}" ::"r"(mbar_smem), "r"(parity) : "memory");
}


// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
inline bool encode_level_map(CUtensorMap* m, const uint8_t* base, const Plan& P, int level, int n_images, int box_w, int box_h) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const LevelGeom& g = P.lv[level];
    const cuuint64_t dims[3] = {(cuuint64_t)g.pitch, (cuuint64_t)g.h, (cuuint64_t)n_images};
    const cuuint64_t strides[2] = {(cuuint64_t)g.pitch, (cuuint64_t)P.pyr_bytes};    // bytes, dims 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u}, estr[3] = {1u, 1u, 1u};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(base) + g.img_off, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, MCV_TMA_L2PROMO, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


}  // namespace mcv
