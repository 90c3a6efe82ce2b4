// Pyramid construction and Gaussian smoothing for a batch of images (sm_100a).
//   k_copy_level0 / k_resize_level : ORBextractor::ComputePyramid (ORBextractor.cc:901-919) — cv::resize INTER_LINEAR u8,
//                                    level l from level l-1, 11-bit fixed-point coefficients (bit-exact model).
//   k_gauss7                       : cv::GaussianBlur 7x7 sigma 2 BORDER_REFLECT_101 (ORBextractor.cc:874-875), Q8 taps
//                                    {18,34,48,56,48,34,18}, (sum + 32768) >> 16.
// Both are HBM/L2-bound byte stencils: 16-byte vector loads/stores, shared-memory tiles, one launch per level over the
// whole batch (grid.z = image).
#include "engine.h"

namespace mcv {

// ---------------------------------------------------------------------------------------------------------
// level 0: copy the caller's image (arbitrary stride) into the pitched pyramid block
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_copy_level0(const uint8_t* __restrict__ src, size_t src_pitch, size_t src_image_stride,
                                                     uint8_t* __restrict__ pyr, int pyr_bytes, int w, int h, int pitch) {
    const int img = blockIdx.z;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (y >= h) return;
    const uint8_t* s = src + (size_t)img * src_image_stride + (size_t)y * src_pitch;
    uint8_t* d = pyr + (size_t)img * pyr_bytes + (size_t)y * pitch;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x0 >= w) return;
    if (x0 + 16 <= w && ((reinterpret_cast<uintptr_t>(s + x0) & 15) == 0)) {
        *reinterpret_cast<uint4*>(d + x0) = __ldg(reinterpret_cast<const uint4*>(s + x0));
    } else {
        for (int x = x0; x < min(x0 + 16, w); ++x) d[x] = s[x];
    }
}

// ---------------------------------------------------------------------------------------------------------
// bilinear resize, 4 output pixels per thread. Tables (host-built, per level): xofs[dw], xa0[dw], xa1[dw], yofs[dh], yb0[dh], yb1[dh]
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize_level(uint8_t* __restrict__ pyr, int pyr_bytes, const int* __restrict__ tab,
                                                      int sw, int sh, int spitch, int soff, int dw, int dh, int dpitch, int doff,
                                                      int area_fast) {
    const int img = blockIdx.z;
    const int dy = blockIdx.y * blockDim.y + threadIdx.y;
    const int dx0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (dy >= dh || dx0 >= dw) return;
    const uint8_t* S = pyr + (size_t)img * pyr_bytes + soff;
    uint8_t* D = pyr + (size_t)img * pyr_bytes + doff + (size_t)dy * dpitch;
    uint32_t packed = 0;
    if (area_fast) {
        const uint8_t* r0 = S + (size_t)(2 * dy) * spitch;
        const uint8_t* r1 = r0 + spitch;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int dx = dx0 + k;
            if (dx < dw) packed |= (uint32_t)((r0[2 * dx] + r0[2 * dx + 1] + r1[2 * dx] + r1[2 * dx + 1] + 2) >> 2) << (8 * k);
        }
    } else {
        const int* xofs = tab;
        const int* xa0 = tab + dw;
        const int* xa1 = tab + 2 * dw;
        const int* yofs = tab + 3 * dw;
        const int* yb0 = yofs + dh;
        const int* yb1 = yofs + 2 * dh;
        int sy0 = yofs[dy], sy1 = sy0 + 1;
        sy0 = min(max(sy0, 0), sh - 1);
        sy1 = min(max(sy1, 0), sh - 1);
        const int b0 = yb0[dy], b1 = yb1[dy];
        const uint8_t* R0 = S + (size_t)sy0 * spitch;
        const uint8_t* R1 = S + (size_t)sy1 * spitch;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int dx = dx0 + k;
            if (dx < dw) {
                const int sx = xofs[dx], sx1 = min(sx + 1, sw - 1);
                const int a0 = xa0[dx], a1 = xa1[dx];
                const int h0 = R0[sx] * a0 + R0[sx1] * a1;
                const int h1 = R1[sx] * a0 + R1[sx1] * a1;
                const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                packed |= (uint32_t)(v & 0xff) << (8 * k);
            }
        }
    }
    if (dx0 + 4 <= dw) {
        *reinterpret_cast<uint32_t*>(D + dx0) = packed;  // dpitch % 16 == 0 and dx0 % 4 == 0
    } else {
        for (int k = 0; dx0 + k < dw; ++k) D[dx0 + k] = (uint8_t)(packed >> (8 * k));
    }
}

int launch_pyramid(const Plan& P, const uint8_t* d_src, size_t src_pitch, size_t src_image_stride, uint8_t* d_pyr, const int* d_tabs,
                   int n_images, cudaStream_t s) {
    int launches = 0;
    {
        const LevelGeom& g = P.lv[0];
        dim3 b(32, 8), grid((g.w + 16 * 32 - 1) / (16 * 32), (g.h + 7) / 8, n_images);
        k_copy_level0<<<grid, b, 0, s>>>(d_src, src_pitch, src_image_stride, d_pyr, P.pyr_bytes, g.w, g.h, g.pitch);
        ++launches;
    }
    for (int l = 1; l < P.n_levels; ++l) {
        const LevelGeom& a = P.lv[l - 1];
        const LevelGeom& g = P.lv[l];
        dim3 b(32, 8), grid((g.w + 4 * 32 - 1) / (4 * 32), (g.h + 7) / 8, n_images);
        k_resize_level<<<grid, b, 0, s>>>(d_pyr, P.pyr_bytes, d_tabs + g.tab_off, a.w, a.h, a.pitch, a.img_off, g.w, g.h, g.pitch,
                                          g.img_off, g.area_fast);
        ++launches;
    }
    return launches;
}

// ---------------------------------------------------------------------------------------------------------
// 7x7 Gaussian, separable, all levels of all images in one launch. Tile = 64 x 32 outputs per CTA.
// ---------------------------------------------------------------------------------------------------------
constexpr int GT_W = 64, GT_H = 32, G_R = 3;

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

struct BlurTiles {  // prefix of tile counts per level so that blockIdx.x -> (level, tile)
    int first_tile[MAX_LEVELS + 1];
    int tiles_x[MAX_LEVELS];
};

__global__ void __launch_bounds__(256) k_gauss7(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur, const __grid_constant__ Plan P,
                                                const __grid_constant__ BlurTiles T) {
    __shared__ uint8_t s_in[GT_H + 2 * G_R][GT_W + 2 * G_R + 2];
    __shared__ uint16_t s_h[GT_H + 2 * G_R][GT_W];
    const int img = blockIdx.y;
    int level = 0;
    while (level + 1 < P.n_levels && (int)blockIdx.x >= T.first_tile[level + 1]) ++level;
    const LevelGeom& g = P.lv[level];
    const int t = blockIdx.x - T.first_tile[level];
    const int tx0 = (t % T.tiles_x[level]) * GT_W, ty0 = (t / T.tiles_x[level]) * GT_H;
    const uint8_t* src = pyr + (size_t)img * P.pyr_bytes + g.img_off;
    uint8_t* dst = blur + (size_t)img * P.pyr_bytes + g.img_off;
    const int tid = threadIdx.x;
    // load tile + halo with reflect-101 at the level border
    for (int i = tid; i < (GT_H + 2 * G_R) * (GT_W + 2 * G_R); i += 256) {
        const int ly = i / (GT_W + 2 * G_R), lx = i % (GT_W + 2 * G_R);
        const int gy = reflect101(ty0 + ly - G_R, g.h), gx = reflect101(tx0 + lx - G_R, g.w);
        s_in[ly][lx] = src[(size_t)gy * g.pitch + gx];
    }
    __syncthreads();
    for (int i = tid; i < (GT_H + 2 * G_R) * GT_W; i += 256) {
        const int ly = i / GT_W, lx = i % GT_W;
        const uint8_t* r = &s_in[ly][lx];
        s_h[ly][lx] = (uint16_t)(18 * (r[0] + r[6]) + 34 * (r[1] + r[5]) + 48 * (r[2] + r[4]) + 56 * r[3]);
    }
    __syncthreads();
    // vertical pass: each thread produces 4 horizontally adjacent pixels of 2 rows
    for (int i = tid; i < GT_H * (GT_W / 4); i += 256) {
        const int ly = i / (GT_W / 4), lx = (i % (GT_W / 4)) * 4;
        const int gy = ty0 + ly, gx = tx0 + lx;
        if (gy >= g.h || gx >= g.w) continue;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t acc = 18u * (s_h[ly][lx + k] + s_h[ly + 6][lx + k]) + 34u * (s_h[ly + 1][lx + k] + s_h[ly + 5][lx + k]) +
                                 48u * (s_h[ly + 2][lx + k] + s_h[ly + 4][lx + k]) + 56u * s_h[ly + 3][lx + k];
            packed |= ((acc + 32768u) >> 16) << (8 * k);
        }
        uint8_t* d = dst + (size_t)gy * g.pitch + gx;
        if (gx + 4 <= g.w) *reinterpret_cast<uint32_t*>(d) = packed;
        else for (int k = 0; gx + k < g.w; ++k) d[k] = (uint8_t)(packed >> (8 * k));
    }
}

int launch_blur(const Plan& P, const uint8_t* d_pyr, uint8_t* d_blur, int n_images, cudaStream_t s) {
    BlurTiles T{};
    int n = 0;
    for (int l = 0; l < P.n_levels; ++l) {
        T.first_tile[l] = n;
        T.tiles_x[l] = (P.lv[l].w + GT_W - 1) / GT_W;
        n += T.tiles_x[l] * ((P.lv[l].h + GT_H - 1) / GT_H);
    }
    for (int l = P.n_levels; l <= MAX_LEVELS; ++l) T.first_tile[l] = n;
    dim3 grid(n, n_images);
    k_gauss7<<<grid, 256, 0, s>>>(d_pyr, d_blur, P, T);
    return 1;
}

}  // namespace mcv
