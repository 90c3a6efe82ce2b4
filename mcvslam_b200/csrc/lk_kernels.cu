// KL_Track's optical flow (src/Frame.cpp:34-76): cv::calcOpticalFlowPyrLK(prev, next, pts, ..., Size(10, 10), maxLevel 1,
// TermCriteria(COUNT + EPS, 10, 0.01), flags 0, minEigThreshold 0.001) for one image pair (sm_100a).
//
// Data-parallel preparation (one thread per pixel):
//   k_lk_level0  — both images into level-0 buffers with the 10-px BORDER_REFLECT_101 ring of buildOpticalFlowPyramid;
//   k_lk_level1  — cv::pyrDown ([1 4 6 4 1]^2, (sum + 128) >> 8, reflect-101) of both images, written with the same ring;
//   k_lk_scharr  — calcScharrDeriv of the previous image's levels, (dx, dy) int16 pairs inside a zero ring.
// Tracking (k_lk_track): one thread per point walks level 1 then level 0 exactly like LKTrackerInvoker::operator(): 14-bit
// fixed-point bilinear window, the 2x2 gradient matrix, up to 10 Newton steps, the L1 residual. The float sums follow the
// SSE2 path of OpenCV's lkpyramid.cpp (four lanes that own pixels x and x + 4 of every window row, pixels 8 and 9 in a scalar
// chain, v_reduce_sum order) — that is what the real cv2 computes (oracle/ora_lk.hpp is pinned bit-exactly against cv2 4.13)
// and therefore what the reference's KL_Track sees. Every float operation is an explicit round-to-nearest intrinsic.
// All points of a pair run concurrently, so the call lasts about as long as one point's serial chain; the preparation
// kernels are pure streaming.
#include "engine.h"

namespace mcv {

constexpr int LK_WIN = 10;
constexpr int LK_W_BITS = 14;

__host__ __device__ inline int lk_reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// Every preparation kernel takes the pair index in blockIdx.z: images of consecutive pairs are `img_stride` bytes apart, their
// workspaces `ws_stride` bytes.
__global__ void k_lk_level0(const uint8_t* __restrict__ prev, const uint8_t* __restrict__ next, int w, int h, int src_pitch, size_t img_stride,
                            uint8_t* __restrict__ I0, uint8_t* __restrict__ J0, int stride, size_t ws_stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= stride) return;
    prev += blockIdx.z * img_stride; next += blockIdx.z * img_stride; I0 += blockIdx.z * ws_stride; J0 += blockIdx.z * ws_stride;
    if (y == h + 2 * LK_WIN) { I0[(size_t)y * stride + x] = 0; J0[(size_t)y * stride + x] = 0; return; }   // spare row (see ora_lk.hpp)
    const int sx = lk_reflect101(x - LK_WIN, w), sy = lk_reflect101(y - LK_WIN, h);
    I0[(size_t)y * stride + x] = prev[(size_t)sy * src_pitch + sx];
    J0[(size_t)y * stride + x] = next[(size_t)sy * src_pitch + sx];
}

__device__ __forceinline__ int lk_pyrdown_px(const uint8_t* __restrict__ src, int w, int h, int pitch, int x, int y) {
    int col[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) col[k] = lk_reflect101(2 * x - 2 + k, w);
    int acc = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint8_t* s = src + (size_t)lk_reflect101(2 * y - 2 + k, h) * pitch;
        const int r = s[col[2]] * 6 + (s[col[1]] + s[col[3]]) * 4 + s[col[0]] + s[col[4]];
        acc += (k == 0 || k == 4) ? r : (k == 2 ? r * 6 : r * 4);
    }
    return (acc + 128) >> 8;
}

__global__ void k_lk_level1(const uint8_t* __restrict__ prev, const uint8_t* __restrict__ next, int w, int h, int src_pitch, size_t img_stride, int w1,
                            int h1, uint8_t* __restrict__ I1, uint8_t* __restrict__ J1, int stride1, size_t ws_stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= stride1) return;
    prev += blockIdx.z * img_stride; next += blockIdx.z * img_stride; I1 += blockIdx.z * ws_stride; J1 += blockIdx.z * ws_stride;
    if (y == h1 + 2 * LK_WIN) { I1[(size_t)y * stride1 + x] = 0; J1[(size_t)y * stride1 + x] = 0; return; }
    const int sx = lk_reflect101(x - LK_WIN, w1), sy = lk_reflect101(y - LK_WIN, h1);
    I1[(size_t)y * stride1 + x] = (uint8_t)lk_pyrdown_px(prev, w, h, src_pitch, sx, sy);
    J1[(size_t)y * stride1 + x] = (uint8_t)lk_pyrdown_px(next, w, h, src_pitch, sx, sy);
}

// img = bordered level (origin at (LK_WIN, LK_WIN)); the reflect-101 ring IS what calcScharrDeriv's border rules read
__global__ void k_lk_scharr(const uint8_t* __restrict__ img, int w, int h, int stride, short2* __restrict__ deriv, size_t ws_stride) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= stride) return;
    img += blockIdx.z * ws_stride; deriv = reinterpret_cast<short2*>(reinterpret_cast<uint8_t*>(deriv) + blockIdx.z * ws_stride);
    const int ix = x - LK_WIN, iy = y - LK_WIN;
    short2 d = make_short2(0, 0);
    if (ix >= 0 && ix < w && iy >= 0 && iy < h) {
        const uint8_t* r0 = img + (size_t)(y - 1) * stride + x;
        const uint8_t* r1 = r0 + stride;
        const uint8_t* r2 = r1 + stride;
        const int t0m = (r0[-1] + r2[-1]) * 3 + r1[-1] * 10, t0p = (r0[1] + r2[1]) * 3 + r1[1] * 10;
        const int t1m = r2[-1] - r0[-1], t1c = r2[0] - r0[0], t1p = r2[1] - r0[1];
        d = make_short2((short)(t0p - t0m), (short)((t1p + t1m) * 3 + t1c * 10));
    }
    deriv[(size_t)y * stride + x] = d;
}

struct LkLevel {
    const uint8_t* I; const uint8_t* J; const short2* D;
    int w, h, stride;
};

__device__ __forceinline__ int lk_floor(float v) { const int i = (int)v; return i - (i > v); }
__device__ __forceinline__ int lk_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
__device__ __forceinline__ void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(na, nb), 16384.f));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, nb), 16384.f));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(na, b), 16384.f));
    w11 = (1 << LK_W_BITS) - w00 - w01 - w10;
}
__device__ __forceinline__ float lk_reduce4(const float* q) { return __fadd_rn(__fadd_rn(q[0], q[2]), __fadd_rn(q[1], q[3])); }

__global__ void __launch_bounds__(64) k_lk_track(LkLevel L0, LkLevel L1, int max_level, size_t ws_stride, const float* __restrict__ pts,
                                                 const int* __restrict__ pair_of, int n, float* __restrict__ next_pts, uint8_t* __restrict__ status_out,
                                                 float* __restrict__ err_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    {   // this point's pair: every level buffer of pair k lies k * ws_stride bytes after pair 0's
        const size_t o = (size_t)(pair_of ? pair_of[i] : 0) * ws_stride;
        L0.I += o; L0.J += o; L0.D = reinterpret_cast<const short2*>(reinterpret_cast<const uint8_t*>(L0.D) + o);
        L1.I += o; L1.J += o; L1.D = reinterpret_cast<const short2*>(reinterpret_cast<const uint8_t*>(L1.D) + o);
    }
    const float px = pts[2 * i], py = pts[2 * i + 1];
    float nx = 0.f, ny = 0.f, err = 0.f;
    int status = 1;
    const float half = 4.5f;                                   // (winSize - 1) * 0.5f
    const float FLT_SCALE = 9.5367431640625e-07f;              // 1.f / (1 << 20)
    short Iwin[LK_WIN * LK_WIN];
    short2 dIwin[LK_WIN * LK_WIN];
    for (int level = max_level; level >= 0; --level) {
        const LkLevel& L = level ? L1 : L0;
        const float lvl_scale = level ? 0.5f : 1.f;
        float prevx = __fmul_rn(px, lvl_scale), prevy = __fmul_rn(py, lvl_scale);
        float nextx, nexty;
        if (level == max_level) { nextx = prevx; nexty = prevy; }
        else { nextx = __fmul_rn(nx, 2.f); nexty = __fmul_rn(ny, 2.f); }
        nx = nextx; ny = nexty;
        prevx = __fsub_rn(prevx, half); prevy = __fsub_rn(prevy, half);
        const int ipx = lk_floor(prevx), ipy = lk_floor(prevy);
        if (ipx < -LK_WIN || ipx >= L.w || ipy < -LK_WIN || ipy >= L.h) {
            if (level == 0) { status = 0; err = 0.f; }
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(prevx, (float)ipx), __fsub_rn(prevy, (float)ipy), w00, w01, w10, w11);
        float iA11 = 0.f, iA12 = 0.f, iA22 = 0.f, qA11[4] = {0.f, 0.f, 0.f, 0.f}, qA12[4] = {0.f, 0.f, 0.f, 0.f}, qA22[4] = {0.f, 0.f, 0.f, 0.f};
        {
            const size_t o = (size_t)(ipy + LK_WIN) * L.stride + (ipx + LK_WIN);
            const uint8_t* src = L.I + o;
            const short2* ds = L.D + o;
            for (int y = 0; y < LK_WIN; ++y, src += L.stride, ds += L.stride) {
#pragma unroll
                for (int x = 0; x < LK_WIN; ++x) {
                    const int ival = lk_descale(src[x] * w00 + src[x + 1] * w01 + src[x + L.stride] * w10 + src[x + L.stride + 1] * w11, LK_W_BITS - 5);
                    const short2 d00 = ds[x], d01 = ds[x + 1], d10 = ds[x + L.stride], d11 = ds[x + L.stride + 1];
                    const int ixval = lk_descale(d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11, LK_W_BITS);
                    const int iyval = lk_descale(d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11, LK_W_BITS);
                    Iwin[y * LK_WIN + x] = (short)ival;
                    dIwin[y * LK_WIN + x] = make_short2((short)ixval, (short)iyval);
                    const float xx = (float)(ixval * ixval), xy = (float)(ixval * iyval), yy = (float)(iyval * iyval);
                    if (x < 8) {
                        qA11[x & 3] = __fadd_rn(qA11[x & 3], xx); qA12[x & 3] = __fadd_rn(qA12[x & 3], xy); qA22[x & 3] = __fadd_rn(qA22[x & 3], yy);
                    } else {
                        iA11 = __fadd_rn(iA11, xx); iA12 = __fadd_rn(iA12, xy); iA22 = __fadd_rn(iA22, yy);
                    }
                }
            }
        }
        iA11 = __fadd_rn(iA11, lk_reduce4(qA11)); iA12 = __fadd_rn(iA12, lk_reduce4(qA12)); iA22 = __fadd_rn(iA22, lk_reduce4(qA22));
        const float A11 = __fmul_rn(iA11, FLT_SCALE), A12 = __fmul_rn(iA12, FLT_SCALE), A22 = __fmul_rn(iA22, FLT_SCALE);
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dA = __fsub_rn(A11, A22);
        const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
                                       (float)(2 * LK_WIN * LK_WIN));
        if (minEig < 0.001f || D < 1.1920928955078125e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nextx = __fsub_rn(nextx, half); nexty = __fsub_rn(nexty, half);
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < 10; ++j) {
            const int inx = lk_floor(nextx), iny = lk_floor(nexty);
            if (inx < -LK_WIN || inx >= L.w || iny < -LK_WIN || iny >= L.h) {
                if (level == 0) status = 0;
                break;
            }
            lk_weights(__fsub_rn(nextx, (float)inx), __fsub_rn(nexty, (float)iny), w00, w01, w10, w11);
            float ib1 = 0.f, ib2 = 0.f, qb1[4] = {0.f, 0.f, 0.f, 0.f}, qb2[4] = {0.f, 0.f, 0.f, 0.f};
            const uint8_t* Jp = L.J + (size_t)(iny + LK_WIN) * L.stride + (inx + LK_WIN);
            for (int y = 0; y < LK_WIN; ++y, Jp += L.stride) {
                int diff[LK_WIN];
#pragma unroll
                for (int x = 0; x < LK_WIN; ++x)
                    diff[x] = lk_descale(Jp[x] * w00 + Jp[x + 1] * w01 + Jp[x + L.stride] * w10 + Jp[x + L.stride + 1] * w11, LK_W_BITS - 5) - Iwin[y * LK_WIN + x];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const short2 da = dIwin[y * LK_WIN + x], db = dIwin[y * LK_WIN + x + 4];
                    qb1[x] = __fadd_rn(qb1[x], (float)(diff[x] * da.x + diff[x + 4] * db.x));
                    qb2[x] = __fadd_rn(qb2[x], (float)(diff[x] * da.y + diff[x + 4] * db.y));
                }
#pragma unroll
                for (int x = 8; x < LK_WIN; ++x) {
                    const short2 da = dIwin[y * LK_WIN + x];
                    ib1 = __fadd_rn(ib1, (float)(diff[x] * da.x));
                    ib2 = __fadd_rn(ib2, (float)(diff[x] * da.y));
                }
            }
            // (qb0 + qb1), pairs interleaved, v_reduce_sum over a vector whose upper half is zero
            ib1 = __fadd_rn(ib1, __fadd_rn(__fadd_rn(__fadd_rn(qb1[0], qb1[2]), 0.f), __fadd_rn(__fadd_rn(qb1[1], qb1[3]), 0.f)));
            ib2 = __fadd_rn(ib2, __fadd_rn(__fadd_rn(__fadd_rn(qb2[0], qb2[2]), 0.f), __fadd_rn(__fadd_rn(qb2[1], qb2[3]), 0.f)));
            const float b1 = __fmul_rn(ib1, FLT_SCALE), b2 = __fmul_rn(ib2, FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nextx = __fadd_rn(nextx, dx); nexty = __fadd_rn(nexty, dy);
            nx = __fadd_rn(nextx, half); ny = __fadd_rn(nexty, half);
            if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= 0.01 * 0.01) break;
            if (j > 0 && (double)fabsf(__fadd_rn(dx, pdx)) < 0.01 && (double)fabsf(__fadd_rn(dy, pdy)) < 0.01) {
                nx = __fsub_rn(nx, __fmul_rn(dx, 0.5f)); ny = __fsub_rn(ny, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {
            const float qx = __fsub_rn(nx, half), qy = __fsub_rn(ny, half);
            const int inx = lk_floor(qx), iny = lk_floor(qy);
            if (inx < -LK_WIN || inx >= L.w || iny < -LK_WIN || iny >= L.h) { status = 0; continue; }
            lk_weights(__fsub_rn(qx, (float)inx), __fsub_rn(qy, (float)iny), w00, w01, w10, w11);
            float errval = 0.f;
            const uint8_t* Jp = L.J + (size_t)(iny + LK_WIN) * L.stride + (inx + LK_WIN);
            for (int y = 0; y < LK_WIN; ++y, Jp += L.stride) {
#pragma unroll
                for (int x = 0; x < LK_WIN; ++x) {
                    const int diff = lk_descale(Jp[x] * w00 + Jp[x + 1] * w01 + Jp[x + L.stride] * w10 + Jp[x + L.stride + 1] * w11, LK_W_BITS - 5) - Iwin[y * LK_WIN + x];
                    errval = __fadd_rn(errval, fabsf((float)diff));
                }
            }
            err = __fdiv_rn(__fmul_rn(errval, 1.f), (float)(32 * LK_WIN * LK_WIN));
        }
    }
    next_pts[2 * i] = nx; next_pts[2 * i + 1] = ny;
    status_out[i] = (uint8_t)status;
    err_out[i] = err;
}

size_t lk_workspace_bytes(int w, int h) {
    const int w1 = (w + 1) / 2, h1 = (h + 1) / 2;
    const size_t l0 = (size_t)(w + 2 * LK_WIN) * (h + 2 * LK_WIN + 1), l1 = (size_t)(w1 + 2 * LK_WIN) * (h1 + 2 * LK_WIN + 1);
    // per level: I, J (u8), D (short2), each rounded up to 256 bytes
    auto r = [](size_t b) { return (b + 255) & ~(size_t)255; };
    return 2 * r(l0) + r(l0 * 4) + 2 * r(l1) + r(l1 * 4);   // a multiple of 256: per-pair workspaces stay aligned
}

// prev / next: n_pairs device images each (rows src_pitch apart, images img_stride apart); d_ws: n_pairs * lk_workspace_bytes(w, h);
// d_pair_of[i] = pair of point i (NULL: one pair). Returns the launch count.
int launch_lk_track(const uint8_t* d_prev, const uint8_t* d_next, int n_pairs, int w, int h, int src_pitch, size_t img_stride, void* d_ws,
                    const float* d_pts, const int* d_pair_of, int n, float* d_next_pts, uint8_t* d_status, float* d_err, cudaStream_t s) {
    const int w1 = (w + 1) / 2, h1 = (h + 1) / 2;
    const int max_level = (w1 <= LK_WIN || h1 <= LK_WIN) ? 0 : 1;     // buildOpticalFlowPyramid stops before a level <= winSize
    const int s0 = w + 2 * LK_WIN, s1 = w1 + 2 * LK_WIN;
    const size_t l0 = (size_t)s0 * (h + 2 * LK_WIN + 1), l1 = (size_t)s1 * (h1 + 2 * LK_WIN + 1);
    auto r = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t ws_stride = lk_workspace_bytes(w, h);
    uint8_t* p = reinterpret_cast<uint8_t*>(d_ws);
    uint8_t* I0 = p; p += r(l0);
    uint8_t* J0 = p; p += r(l0);
    short2* D0 = reinterpret_cast<short2*>(p); p += r(l0 * 4);
    uint8_t* I1 = p; p += r(l1);
    uint8_t* J1 = p; p += r(l1);
    short2* D1 = reinterpret_cast<short2*>(p);
    int launches = 0;
    k_lk_level0<<<dim3((s0 + 255) / 256, h + 2 * LK_WIN + 1, n_pairs), 256, 0, s>>>(d_prev, d_next, w, h, src_pitch, img_stride, I0, J0, s0, ws_stride); ++launches;
    k_lk_scharr<<<dim3((s0 + 255) / 256, h + 2 * LK_WIN + 1, n_pairs), 256, 0, s>>>(I0, w, h, s0, D0, ws_stride); ++launches;
    if (max_level == 1) {
        k_lk_level1<<<dim3((s1 + 255) / 256, h1 + 2 * LK_WIN + 1, n_pairs), 256, 0, s>>>(d_prev, d_next, w, h, src_pitch, img_stride, w1, h1, I1, J1, s1, ws_stride); ++launches;
        k_lk_scharr<<<dim3((s1 + 255) / 256, h1 + 2 * LK_WIN + 1, n_pairs), 256, 0, s>>>(I1, w1, h1, s1, D1, ws_stride); ++launches;
    }
    if (n > 0) {
        const LkLevel L0{I0, J0, D0, w, h, s0}, L1{I1, J1, D1, w1, h1, s1};
        k_lk_track<<<(n + 63) / 64, 64, 0, s>>>(L0, L1, max_level, ws_stride, d_pts, d_pair_of, n, d_next_pts, d_status, d_err); ++launches;
    }
    return launches;
}

}  // namespace mcv
