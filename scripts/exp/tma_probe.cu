// Probe: 3-D u8 tensor map (pitch, h, n_images), unaligned box start, maps passed inside a __grid_constant__ struct.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
struct Maps { CUtensorMap m[4]; };
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned mbar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(mbar) : "memory");
}
__global__ void k(const __grid_constant__ Maps maps, int which, int x, int y, int z, uint8_t* out, int mode) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned mb = (unsigned)__cvta_generic_to_shared(&bar);
    const unsigned bs = (unsigned)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(1) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.release.cta.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(mb), "r"(48 * 37) : "memory");
        if (mode == 0) tma_load_3d(bs, &maps.m[which], x, y, z, mb);
    }
    if (mode == 0) {
        asm volatile(
            "{\n\t.reg .pred p;\nW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\nD_%=:\n\t}" ::"r"(mb), "r"(0) : "memory");
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 48 * 37; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int pitch = 640, h = 480, n = 3;
    std::vector<uint8_t> img((size_t)pitch * h * n);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *o;
    cudaMalloc(&d, img.size()); cudaMalloc(&o, 48 * 37);
    cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    Maps maps; memset(&maps, 0, sizeof(maps));
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)h, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
    const cuuint32_t box[3] = {48, 37, 1}, es[3] = {1, 1, 1};
    for (int i = 0; i < 4; ++i) {
        CUresult r = fn(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d -> %d\n", i, (int)r);
    }
    const int cases[][4] = {{0, 0, 0, 0}, {1, 16, 5, 1}, {2, 17, 5, 2}, {3, 601, 450, 1}, {1, 3, 7, 0}};
    for (auto& c : cases) {
        cudaMemset(o, 0xee, 48 * 37);
        k<<<1, 128, 48 * 37 + 128>>>(maps, c[0], c[1], c[2], c[3], o, 0);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint8_t> r(48 * 37);
        cudaMemcpy(r.data(), o, r.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < 37; ++yy) for (int xx = 0; xx < 48; ++xx) {
            const int gx = c[1] + xx, gy = c[2] + yy;
            const uint8_t want = (gx < pitch && gy < h) ? img[(size_t)c[3] * pitch * h + (size_t)gy * pitch + gx] : 0;
            bad += r[yy * 48 + xx] != want;
        }
        printf("case map %d at (%d,%d,%d): %s, mismatches %d\n", c[0], c[1], c[2], c[3], cudaGetErrorString(e), bad);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
