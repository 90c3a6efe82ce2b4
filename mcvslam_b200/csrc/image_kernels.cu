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
// level 0 from an interleaved BGR image: cv::cvtColor(COLOR_BGR2GRAY) of System::Track (src/System.cpp:60-64) fused into
// the level-0 write. OpenCV's 8-bit path is fixed point: (B * 3735 + G * 19235 + R * 9798 + (1 << 14)) >> 15 (pinned against
// cv2 4.13, tests/golden/bgr_golden.npz). 4 pixels per thread: three aligned 32-bit loads, one 32-bit store.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned bgr_gray(unsigned b, unsigned g, unsigned r) { return (b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15; }

__global__ void __launch_bounds__(256) k_gray_level0(const uint8_t* __restrict__ src, size_t src_pitch, size_t src_image_stride,
                                                     uint8_t* __restrict__ pyr, int pyr_bytes, int w, int h, int pitch) {
    const int img = blockIdx.z;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (y >= h) return;
    const uint8_t* s = src + (size_t)img * src_image_stride + (size_t)y * src_pitch;
    uint8_t* d = pyr + (size_t)img * pyr_bytes + (size_t)y * pitch;
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x0 >= w) return;
    const uint8_t* p = s + 3 * (size_t)x0;
    if (x0 + 4 <= w && ((reinterpret_cast<uintptr_t>(p) & 3) == 0)) {
        const unsigned w0 = __ldg(reinterpret_cast<const unsigned*>(p)), w1 = __ldg(reinterpret_cast<const unsigned*>(p + 4)),
                       w2 = __ldg(reinterpret_cast<const unsigned*>(p + 8));   // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
        const unsigned g0 = bgr_gray(w0 & 0xff, (w0 >> 8) & 0xff, (w0 >> 16) & 0xff);
        const unsigned g1 = bgr_gray(w0 >> 24, w1 & 0xff, (w1 >> 8) & 0xff);
        const unsigned g2 = bgr_gray((w1 >> 16) & 0xff, w1 >> 24, w2 & 0xff);
        const unsigned g3 = bgr_gray((w2 >> 8) & 0xff, (w2 >> 16) & 0xff, w2 >> 24);
        *reinterpret_cast<unsigned*>(d + x0) = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);   // level rows are 16-byte aligned
    } else {
        for (int x = x0; x < min(x0 + 4, w); ++x) d[x] = (uint8_t)bgr_gray(s[3 * x], s[3 * x + 1], s[3 * x + 2]);
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

// ---------------------------------------------------------------------------------------------------------
// bilinear resize, register-marching form for the pyramid's usual geometry (1 < scale <= 2, strictly increasing source
// rows; the host checks both on the tables and otherwise uses k_resize_level).
//
// A warp owns a 128-px wide strip of the destination (lane = 4 output pixels) and marches down RS_ROWS output rows. All
// x-geometry is row-invariant, so each lane builds it ONCE: the 8-byte source window that holds S[sx], S[sx+1] of its four
// pixels starts at byte sx(px0); it is assembled from three aligned 32-bit loads with two funnel shifts, and two
// byte-permutes put (S0, S1) pairs where DP2A wants them — the horizontal pass H = S0*a0 + S1*a1 is one DP2A per pixel.
// The horizontal result of a source row is kept in registers and reused by the next output row (each source row feeds
// ~1.7 output rows at scale 1.2). Vertical pass per pixel: 2 IMAD.HI + IADD3 + SHF, i.e. exactly
// (((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2 of cv::resize's 11-bit fixed-point path.
// ---------------------------------------------------------------------------------------------------------
#ifndef MCV_RS_ROWS
#define MCV_RS_ROWS 16
#endif
#ifndef MCV_RS_WARPS
#define MCV_RS_WARPS 1
#endif
constexpr int RS_ROWS = MCV_RS_ROWS;   // output rows per warp (<= 32: lane j holds row j's coefficients)
constexpr int RS_WARPS = MCV_RS_WARPS;
constexpr int RS_PREF = 4;    // source rows in flight per lane
// Tried and dropped (bit-exact, B200): the leftover columns of a level folded as in k_gauss7 (32 / width bands per warp: 433 -> 373
// warps per image). Each lane group then needs its own row schedule, so the row coefficients come from the tables per lane (three
// dependent loads per output row instead of three shuffles from "lane j") and the emission loop diverges between groups: pyramid
// stage 0.394 -> 0.435 ms.

__global__ void __launch_bounds__(32 * RS_WARPS) k_resize_march(uint8_t* __restrict__ pyr, int pyr_bytes, const int* __restrict__ tab,
                                                                 int sw, int sh, int spitch, int soff, int dw, int dh, int dpitch, int doff,
                                                                 int strips_x, int n_strips) {
    const int img = blockIdx.y, lane = threadIdx.x & 31;
    const int sid = blockIdx.x * RS_WARPS + (threadIdx.x >> 5);
    if (sid >= n_strips) return;
    const int dx0 = (sid % strips_x) * 128 + lane * 4, dy0 = (sid / strips_x) * RS_ROWS;
    const uint8_t* S = pyr + (size_t)img * pyr_bytes + soff;
    uint8_t* D = pyr + (size_t)img * pyr_bytes + doff;
    const int* xofs = tab;
    const int* xa0 = tab + dw;
    const int* xa1 = tab + 2 * dw;
    const int* yofs = tab + 3 * dw;
    const int* yb0 = yofs + dh;
    const int* yb1 = yofs + 2 * dh;
    const bool active = dx0 < dw;
    // row-invariant x geometry of this lane
    unsigned coef[4], sel01 = 0, sel23 = 0;
    int wb = 0, shift = 0;
    {
        int sx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int dx = min(dx0 + k, dw - 1);
            sx[k] = active ? xofs[dx] : 0;
            coef[k] = active ? ((unsigned)xa0[dx] | ((unsigned)xa1[dx] << 16)) : 0u;
        }
        wb = sx[0] & ~3; shift = (sx[0] & 3) * 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const unsigned o0 = (unsigned)(sx[k] - sx[0]), o1 = (unsigned)(min(sx[k] + 1, sw - 1) - sx[0]);
            const unsigned pair = (o0 & 7u) | ((o1 & 7u) << 4);
            if (k < 2) sel01 |= pair << (8 * k); else sel23 |= pair << (8 * (k - 2));
        }
    }
    // per-row coefficients of this strip: lane j holds output row dy0 + j
    int my_rb = -1, my_same = 0, my_sy = 0;
    unsigned my_b0 = 0, my_b1 = 0;
    if (dy0 + lane < dh) {
        my_sy = yofs[dy0 + lane]; my_b0 = (unsigned)yb0[dy0 + lane] << 16; my_b1 = (unsigned)yb1[dy0 + lane] << 16;
        const int ra = min(max(my_sy, 0), sh - 1);
        my_rb = min(max(my_sy + 1, 0), sh - 1);
        my_same = ra == my_rb;                                      // both source rows clamp to the same row (image edge)
    }
    auto hpass = [&](const unsigned w[3], unsigned h[4]) {         // horizontal pass (>> 4 applied)
        const unsigned lo = __funnelshift_r(w[0], w[1], shift), hi = __funnelshift_r(w[1], w[2], shift);
        const unsigned p01 = __byte_perm(lo, hi, sel01), p23 = __byte_perm(lo, hi, sel23);
        h[0] = __dp2a_lo(coef[0], p01, 0u) >> 4; h[1] = __dp2a_hi(coef[1], p01, 0u) >> 4;
        h[2] = __dp2a_lo(coef[2], p23, 0u) >> 4; h[3] = __dp2a_hi(coef[3], p23, 0u) >> 4;
    };

    // The strip's source rows are consecutive: walk them once with RS_PREF rows in flight (statically indexed ring, no
    // register rotation: the two horizontal results alternate between hA and hB), and emit every output row as soon as its
    // lower source row (rb) has been filtered. The three loads of a row are p[0], p[4], p[8] off one running pointer; words
    // past the row end are never selected (the pyramid allocation carries 256 spare bytes for the very last row).
    const int rows = min(RS_ROWS, dh - dy0);
    const int r_begin = min(max(__shfl_sync(0xffffffffu, my_sy, 0), 0), sh - 1);
    const int r_end = __shfl_sync(0xffffffffu, my_rb, rows - 1);
    const uint8_t* sp = S + (size_t)r_begin * spitch + wb;          // row r_begin; advanced by RS_PREF rows per round
    uint8_t* dp = D + (size_t)dy0 * dpitch + dx0;
    unsigned raw[RS_PREF][3];
#pragma unroll
    for (int u = 0; u < RS_PREF; ++u) {
        const bool ld = active && r_begin + u <= r_end;
        const uint8_t* p = sp + u * spitch;
        raw[u][0] = ld ? *reinterpret_cast<const unsigned*>(p) : 0u;
        raw[u][1] = ld ? *reinterpret_cast<const unsigned*>(p + 4) : 0u;
        raw[u][2] = ld ? *reinterpret_cast<const unsigned*>(p + 8) : 0u;
    }
    unsigned hA[4] = {0, 0, 0, 0}, hB[4] = {0, 0, 0, 0};
    int j = 0, rb_next = __shfl_sync(0xffffffffu, my_rb, 0);
#pragma unroll 1
    for (int r = r_begin; r <= r_end; r += RS_PREF) {
        sp += RS_PREF * spitch;
#pragma unroll
        for (int u = 0; u < RS_PREF; ++u) {
            const int rr = r + u;
            const unsigned cur[3] = {raw[u][0], raw[u][1], raw[u][2]};
            {
                const bool ld = active && rr + RS_PREF <= r_end;
                const uint8_t* p = sp + u * spitch;
                raw[u][0] = ld ? *reinterpret_cast<const unsigned*>(p) : 0u;
                raw[u][1] = ld ? *reinterpret_cast<const unsigned*>(p + 4) : 0u;
                raw[u][2] = ld ? *reinterpret_cast<const unsigned*>(p + 8) : 0u;
            }
            unsigned* hc = (u & 1) ? hB : hA;                       // resolved at compile time
            unsigned* hp = (u & 1) ? hA : hB;
            hpass(cur, hc);
            while (rb_next == rr) {
                const unsigned b0 = __shfl_sync(0xffffffffu, my_b0, j), b1 = __shfl_sync(0xffffffffu, my_b1, j);
                const bool same = __shfl_sync(0xffffffffu, my_same, j) != 0;
                unsigned v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = (__umulhi(b0, same ? hc[k] : hp[k]) + __umulhi(b1, hc[k]) + 2u) >> 2;
                const unsigned packed = __byte_perm(__byte_perm(v[0], v[1], 0x0040), __byte_perm(v[2], v[3], 0x0040), 0x5410);
                if (dx0 + 4 <= dw) *reinterpret_cast<unsigned*>(dp) = packed;
                else if (active) for (int k = 0; dx0 + k < dw; ++k) dp[k] = (uint8_t)(packed >> (8 * k));
                dp += dpitch;
                ++j;
                rb_next = j < rows ? __shfl_sync(0xffffffffu, my_rb, j) : -1;
            }
        }
    }
}

int launch_pyramid(const Plan& P, const uint8_t* d_src, size_t src_pitch, size_t src_image_stride, int src_channels, uint8_t* d_pyr,
                   const int* d_tabs, int n_images, cudaStream_t s) {
    int launches = 0;
    {
        const LevelGeom& g = P.lv[0];
        if (src_channels == 3) {
            dim3 b(32, 8), grid((g.w + 4 * 32 - 1) / (4 * 32), (g.h + 7) / 8, n_images);
            k_gray_level0<<<grid, b, 0, s>>>(d_src, src_pitch, src_image_stride, d_pyr, P.pyr_bytes, g.w, g.h, g.pitch);
        } else {
            dim3 b(32, 8), grid((g.w + 16 * 32 - 1) / (16 * 32), (g.h + 7) / 8, n_images);
            k_copy_level0<<<grid, b, 0, s>>>(d_src, src_pitch, src_image_stride, d_pyr, P.pyr_bytes, g.w, g.h, g.pitch);
        }
        ++launches;
    }
    for (int l = 1; l < P.n_levels; ++l) {
        const LevelGeom& a = P.lv[l - 1];
        const LevelGeom& g = P.lv[l];
        if (g.march_ok) {
            const int strips_x = (g.w + 127) / 128, n_strips = strips_x * ((g.h + RS_ROWS - 1) / RS_ROWS);
            k_resize_march<<<dim3((n_strips + RS_WARPS - 1) / RS_WARPS, n_images), 32 * RS_WARPS, 0, s>>>(
                d_pyr, P.pyr_bytes, d_tabs + g.tab_off, a.w, a.h, a.pitch, a.img_off, g.w, g.h, g.pitch, g.img_off, strips_x, n_strips);
        } else {
            dim3 b(32, 8), grid((g.w + 4 * 32 - 1) / (4 * 32), (g.h + 7) / 8, n_images);
            k_resize_level<<<grid, b, 0, s>>>(d_pyr, P.pyr_bytes, d_tabs + g.tab_off, a.w, a.h, a.pitch, a.img_off, g.w, g.h, g.pitch,
                                              g.img_off, g.area_fast);
        }
        ++launches;
    }
    return launches;
}

// ---------------------------------------------------------------------------------------------------------
// 7x7 Gaussian (Q8 taps 18,34,48,56,48,34,18; two exact integer passes), all levels of all images in one launch.
//
// Register-marching separable filter, no shared memory: a warp owns a 128-px wide strip (lane = one aligned 4-px word) and
// marches down GB_ROWS rows. Per row every lane loads ONE coalesced 32-bit word, gets its left/right neighbour words by
// warp shuffle, forms the 7-tap horizontal sums with two dp4a per pixel, and keeps the last 7 rows of sums in registers
// (statically indexed ring, two consecutive rows' 16-bit sums per register); the vertical pass is 3 DP2A + 1 IMAD per pixel on
// that ring. HBM traffic = 1 read + 1 write per pixel plus the 6-row halo (19 %).
// ---------------------------------------------------------------------------------------------------------
#ifndef MCV_GB_ROWS
#define MCV_GB_ROWS 50
#endif
#ifndef MCV_GB_WARPS
#define MCV_GB_WARPS 1
#endif
constexpr int GB_ROWS = MCV_GB_ROWS;              // output rows per warp; GB_ROWS + 6 is a multiple of the 7-row ring
constexpr int GB_WARPS = MCV_GB_WARPS;

__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

// One warp per CTA, 24 resident CTAs per SM = 80 registers (one 4-byte spill outside the loop), 50-row strips (12 % halo rows).
// B200, 384 images, with the DP2A vertical pass: 4 warps x 36 rows, no register target (85 registers) 0.394 ms; 4 x 6 CTAs
// (80 registers) 0.368; x 7 (72) 0.477; 1 warp x 24 CTAs 0.364, x 28 0.390; 1 x 24 with 29 / 50 / 64 / 78 / 120 rows:
// 0.372 / 0.343 / 0.358 / 0.357 / 0.352. Leftover columns folded (several bands per warp): 0.343 -> 0.318 ms.
#ifndef GB_MINB
#define GB_MINB 24
#endif
__global__ void __launch_bounds__(32 * GB_WARPS, GB_MINB) k_gauss7(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur,
                                                           const __grid_constant__ Plan P, const __grid_constant__ StripTable T) {
    const int img = blockIdx.y, lane = threadIdx.x & 31;
    const int sid = blockIdx.x * GB_WARPS + (threadIdx.x >> 5);
    if (sid >= T.first[P.n_levels]) return;
    int level = 0;
    while (level + 1 < P.n_levels && sid >= T.first[level + 1]) ++level;
    const LevelGeom& g = P.lv[level];
    const int t = sid - T.first[level];
    const int w = g.w, h = g.h, pitch = g.pitch;
    // Strips of a level: first its FULL 128-px strips (strips_x per band of GB_ROWS rows), then the columns that are left over
    // (w mod 128, rl lanes wide) FOLDED: a warp takes 32 / rl bands of them side by side, so that a level whose width is not a
    // multiple of 128 does not leave the rest of a warp idle on every band (levels 1-7 of 640x480: 177 -> 162 warps per image).
    const int nxf = T.strips_x[level], n_bands = (h + GB_ROWS - 1) / GB_ROWS, n_full = nxf * n_bands;
    int x0, y0;
    bool first_l, right_l;                                  // lane fetches the outer neighbour word on its left / right itself
    if (t < n_full) {
        x0 = (t % nxf) * 128 + lane * 4; y0 = (t / nxf) * GB_ROWS;
        first_l = lane == 0; right_l = lane == 31;
    } else {
        const int rl = (w - nxf * 128 + 3) >> 2, fold = 32 / rl, grp = lane / rl, lg = lane - grp * rl;
        const int band = (t - n_full) * fold + grp;
        const bool mine = grp < fold && band < n_bands;
        x0 = mine ? nxf * 128 + lg * 4 : w + 64; y0 = mine ? band * GB_ROWS : 0;
        first_l = lg == 0; right_l = false;                 // a group's last lane is at the image's right edge: reflected, no neighbour
    }
    const uint8_t* src = pyr + (size_t)img * P.pyr_bytes + g.img_off;
    uint8_t* dst = blur + (size_t)img * P.pyr_bytes + g.img_off + x0;
    const bool active = x0 < w;
    const int n_valid = min(4, w - x0);                     // pixels of this lane's word inside the image (<= 0: none)
    // the first / last lane of a strip fetch the strip's outer neighbour words themselves (one predicated load per row)
    const int ex = first_l ? x0 - 4 : x0 + 4;
    const bool edge = active && ((first_l && x0 > 0) || (right_l && x0 + 4 < pitch));
    const uint32_t TA = 18u | (34u << 8) | (48u << 16) | (56u << 24), TB = 48u | (34u << 8) | (18u << 16);
    // BORDER_REFLECT_101 in x without branches in the row loop: per-lane byte-permute selectors, built once.
    // Window bytes b[-4..7] = (w0, w1, w2). Left edge (x0 == 0): b[-i] = b[i]. Right edge at e = w - 1 - x0: b[i] = b[2e - i], i > e.
    const int e = w - 1 - x0;
    uint32_t sel0 = x0 == 0 ? 0x5670u : 0x3210u, sel1 = 0x7654u, sel2 = 0x7654u;   // identity on (w0,w1) -> w0 / w1, on (w1,w2) -> w2
    const bool w2_from_w0w1 = e <= 2;                       // sources of the reflected w2 bytes then lie in (w0, w1)
    if (active && e < 7) {
        sel1 = 0; sel2 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int s1 = i > e ? 2 * e - i : i;           // source pixel index (window coordinates) of b'[i]
            sel1 |= (uint32_t)((s1 + 4) & 7) << (4 * i);
            const int i2 = i + 4, s2 = i2 > e ? 2 * e - i2 : i2;
            sel2 |= (uint32_t)((w2_from_w0w1 ? s2 + 4 : s2) & 7) << (4 * i);
        }
    }

    auto row_ptr = [&](int r) {                             // reflected input row y0 - 3 + r, clamped for rows that feed nothing
        int y = y0 - 3 + r;
        y = y < 0 ? -y : y;
        y = y >= h ? 2 * h - 2 - y : y;
        return src + min(max(y, 0), h - 1) * pitch;
    };
    // 7-deep software prefetch of the row words (statically indexed, like the ring)
    uint32_t pw[7], pe[7];
#pragma unroll
    for (int d = 0; d < 7; ++d) {
        const uint8_t* row = row_ptr(d);
        pw[d] = active ? __ldg(reinterpret_cast<const uint32_t*>(row + x0)) : 0u;
        pe[d] = edge ? __ldg(reinterpret_cast<const uint32_t*>(row + ex)) : 0u;
    }
    // ring[j][k]: the horizontal sums (< 2^16) of the row ring slot j was written for in the HIGH half and of the row before it in
    // the LOW half, so that the vertical pass is three DP2A (two rows each) + one IMAD per pixel instead of 3 IADD + 4 IMAD
    uint32_t ring[7][4];
#pragma unroll
    for (int j = 0; j < 7; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) ring[j][k] = 0;
    const uint32_t VA = 18u | (34u << 8) | (48u << 16) | (56u << 24), VB = 48u | (34u << 8);

#pragma unroll 1
    for (int rb = 0; rb < GB_ROWS + 6; rb += 7) {
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const int r = rb + j;                           // input row y0 - 3 + r
            const uint32_t w1r = pw[j], we = pe[j];
            {
                const uint8_t* row = row_ptr(r + 7);
                pw[j] = active ? __ldg(reinterpret_cast<const uint32_t*>(row + x0)) : 0u;
                pe[j] = edge ? __ldg(reinterpret_cast<const uint32_t*>(row + ex)) : 0u;
            }
            uint32_t w0r = __shfl_up_sync(0xffffffffu, w1r, 1), w2r = __shfl_down_sync(0xffffffffu, w1r, 1);
            w0r = first_l ? we : w0r;
            w2r = right_l ? we : w2r;
            const uint32_t w0 = __byte_perm(w0r, w1r, sel0), w1 = __byte_perm(w0r, w1r, sel1);
            const uint32_t w2 = __byte_perm(w2_from_w0w1 ? w0r : w1r, w2_from_w0w1 ? w1r : w2r, sel2);
            // horizontal 7-tap: px k uses bytes [k-3, k] of (w0:w1) and [k+1, k+4] of (w1:w2)
            uint32_t hs[4];
            hs[0] = __dp4a(__funnelshift_r(w0, w1, 8), TA, __dp4a(__funnelshift_r(w1, w2, 8), TB, 0u));
            hs[1] = __dp4a(__funnelshift_r(w0, w1, 16), TA, __dp4a(__funnelshift_r(w1, w2, 16), TB, 0u));
            hs[2] = __dp4a(__funnelshift_r(w0, w1, 24), TA, __dp4a(__funnelshift_r(w1, w2, 24), TB, 0u));
            hs[3] = __dp4a(w1, TA, __dp4a(w2, TB, 0u));
#pragma unroll
            for (int k = 0; k < 4; ++k) ring[j][k] = __byte_perm(ring[(j + 6) % 7][k], hs[k], 0x5432);   // (previous row, this row)
            const int y = y0 + r - 6;                       // output row whose window ends at input row r
            if (r >= 6 && y < h) {
                uint32_t acc[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)                 // rows r-6 .. r: 18 34 | 48 56 | 48 34 | 18
                    acc[k] = __dp2a_lo(ring[(j + 6) % 7][k], VB, __dp2a_hi(ring[(j + 4) % 7][k], VA, __dp2a_lo(ring[(j + 2) % 7][k], VA, hs[k] * 18u + 32768u)));
                // byte 2 of each accumulator -> one word
                const uint32_t o = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
                uint8_t* d = dst + y * pitch;
                if (n_valid == 4) *reinterpret_cast<uint32_t*>(d) = o;
                else if (n_valid > 0) {
                    d[0] = (uint8_t)o;
                    if (n_valid > 1) d[1] = (uint8_t)(o >> 8);
                    if (n_valid > 2) d[2] = (uint8_t)(o >> 16);
                }
            }
        }
    }
}

int launch_blur(const Plan& P, const uint8_t* d_pyr, uint8_t* d_blur, int n_images, cudaStream_t s) {
    StripTable T{};
    int n = 0;
    for (int l = 0; l < P.n_levels; ++l) {
        T.first[l] = n;
        T.strips_x[l] = P.lv[l].w / 128;                    // FULL strips per band; the remaining columns are folded (k_gauss7)
        const int n_bands = (P.lv[l].h + GB_ROWS - 1) / GB_ROWS, rl = (P.lv[l].w - T.strips_x[l] * 128 + 3) / 4;
        n += T.strips_x[l] * n_bands + (rl ? (n_bands + 32 / rl - 1) / (32 / rl) : 0);
    }
    for (int l = P.n_levels; l <= MAX_LEVELS; ++l) T.first[l] = n;
    dim3 grid((n + GB_WARPS - 1) / GB_WARPS, n_images);
    k_gauss7<<<grid, 32 * GB_WARPS, 0, s>>>(d_pyr, d_blur, P, T);
    return 1;
}

}  // namespace mcv
