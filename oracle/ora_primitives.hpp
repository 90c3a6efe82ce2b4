// ORACLE — TEST INFRASTRUCTURE ONLY. Not product code: nothing under mcvslam_b200/ may include, link or call this.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the checker.
//
// CPU restatements of the OpenCV 4 primitives the reference's hot path calls (OpenCV is a third-party dependency
// of Sologala/MCVSLAM that is NOT vendored in /root/reference and not pinned: cmake/FindOpenCV.cmake:10
// `find_package(OpenCV 4 REQUIRED)`). Each model below restates the published OpenCV 4.x algorithm and is pinned
// against the in-container cv2 4.13.0 by tests/test_oracle_vs_cv2.py and by the committed fixtures in tests/golden/.
//
// Parity status: the reference ships no golden vectors / assertions for this path, so the pin is
// cv2 4.13.0 (the dependency that owns the arithmetic) — see DESIGN.md "Oracle".
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace ora {

// cvRound(float/double): SSE2 cvtss2si / cvtsd2si == round-half-to-even in the default rounding mode.
static inline int cv_round(float v) { return (int)lrintf(v); }
static inline int cv_round(double v) { return (int)lrint(v); }
static inline int cv_floor(double v) { int i = (int)v; return i - (i > v); }

struct Img {  // non-owning u8 view, like a cv::Mat header (CV_8UC1)
    const uint8_t* data; int w, h; size_t stride;
    const uint8_t* row(int y) const { return data + (size_t)y * stride; }
};

// ---------------------------------------------------------------------------------------------------------
// cv::resize(src, dst, sz, 0, 0, INTER_LINEAR) for CV_8UC1 — called at ORBextractor.cc:911.
// OpenCV imgproc/resize.cpp: fixed-point bilinear, INTER_RESIZE_COEF_BITS=11; HResizeLinear<uchar,int,short>
// produces int32 rows, VResizeLinear<uchar,int,short> combines with the ">>4, *beta >>16, +2 >>2" sequence.
// The exact-2x decimation special case (INTER_LINEAR -> INTER_AREA fast path) is included.
// ---------------------------------------------------------------------------------------------------------
static inline void resize_linear_u8(const Img& s, uint8_t* dst, int dw, int dh, size_t dstride) {
    const int sw = s.w, sh = s.h;
    const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    const double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    const int iscale_x = (int)lrint(scale_x), iscale_y = (int)lrint(scale_y);  // saturate_cast<int>(double)
    const bool is_area_fast = std::abs(scale_x - iscale_x) < DBL_EPSILON && std::abs(scale_y - iscale_y) < DBL_EPSILON;
    if (is_area_fast && iscale_x == 2 && iscale_y == 2) {
        // ResizeAreaFastVec 2x2 box: (a+b+c+d+2)>>2
        for (int y = 0; y < dh; ++y) {
            const uint8_t* r0 = s.row(2 * y); const uint8_t* r1 = s.row(2 * y + 1);
            for (int x = 0; x < dw; ++x) dst[y * dstride + x] = (uint8_t)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2);
        }
        return;
    }
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> ialpha(2 * dw), ibeta(2 * dh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx] = (short)cv_round((1.f - fx) * 2048.f);
        ialpha[2 * dx + 1] = (short)cv_round(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        ibeta[2 * dy] = (short)cv_round((1.f - fy) * 2048.f);
        ibeta[2 * dy + 1] = (short)cv_round(fy * 2048.f);
    }
    std::vector<int> row0(dw), row1(dw);
    auto hresize = [&](int sy, std::vector<int>& out) {
        sy = sy < 0 ? 0 : (sy >= sh ? sh - 1 : sy);  // clip(sy, 0, ssize.height)
        const uint8_t* S = s.row(sy);
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx];
            int sx1 = sx + 1 < sw ? sx + 1 : sw - 1;  // coefficient is 0 there (dx >= xmax path: S[sx]*ONE)
            out[dx] = S[sx] * ialpha[2 * dx] + S[sx1] * ialpha[2 * dx + 1];
        }
    };
    for (int dy = 0; dy < dh; ++dy) {
        hresize(yofs[dy], row0);
        hresize(yofs[dy] + 1, row1);
        const int b0 = ibeta[2 * dy], b1 = ibeta[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
        }
    }
}

static inline int reflect101(int p, int len) {  // cv::borderInterpolate(p, len, BORDER_REFLECT_101)
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
    return p;
}

// ---------------------------------------------------------------------------------------------------------
// cv::GaussianBlur(img, img, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 — ORBextractor.cc:875.
// OpenCV 4 smooth.dispatch.cpp: 8-bit path uses ufixedpoint16 kernels from getGaussianKernelFixedPoint_ED; for
// ksize 7 / sigma 2 the Q8 taps are {18,34,48,56,48,34,18} (sum 256). Horizontal pass keeps Q8.8 (no rounding,
// max 255*256 fits u16), vertical pass accumulates u32 and rounds: (sum + 32768) >> 16.
// ---------------------------------------------------------------------------------------------------------
static const int GAUSS7_Q8[7] = {18, 34, 48, 56, 48, 34, 18};

static inline void gauss7_u8(const Img& s, uint8_t* dst, size_t dstride) {
    const int w = s.w, h = s.h;
    std::vector<uint16_t> tmp((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* S = s.row(y);
        for (int x = 0; x < w; ++x) {
            int acc = 0;
            for (int k = -3; k <= 3; ++k) acc += GAUSS7_Q8[k + 3] * S[reflect101(x + k, w)];
            tmp[(size_t)y * w + x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* R[7];
        for (int k = -3; k <= 3; ++k) R[k + 3] = &tmp[(size_t)reflect101(y + k, h) * w];
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; ++k) acc += (uint32_t)GAUSS7_Q8[k] * R[k][x];
            dst[(size_t)y * dstride + x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// cv::fastAtan2(y, x) — ORBextractor.cc:97. OpenCV core/mathfuncs_core.simd.hpp atan_f32: 7th-order odd minimax
// polynomial in degrees, separate f32 multiplies and adds (no FMA contraction in the shipped baseline build).
// ---------------------------------------------------------------------------------------------------------
static inline float fast_atan2(float y, float x) {
    static const float scale = (float)(180.0 / 3.14159265358979323846);
    static const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                       p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    volatile float t;  // force every intermediate through f32 (belt and braces on x86-64/SSE: already true)
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1; t = t * c;
        a = t;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1; t = t * c;
        a = 90.f - t;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------------------------
// cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16 — ORBextractor.cc:619,622.
// OpenCV features2d/fast.cpp FAST_t<16> + fast_score.cpp cornerScore<16>:
//   * Bresenham circle of 16, corner iff 9 contiguous pixels all brighter than v+t or all darker than v-t;
//   * score = (largest t' for which the pixel is still a corner) = max(max_arc min(v-p), max_arc min(p-v)) - 1;
//   * only rows/cols [3, n-3) of the given (sub-)image are examined; non-corners score 0;
//   * NMS keeps a corner iff its score is strictly greater than all 8 neighbours' scores;
//   * keypoints come out row-major with KeyPoint(x, y, 7.f, -1, score) (octave 0, class_id -1).
// ---------------------------------------------------------------------------------------------------------
static const int FAST_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int FAST_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// Returns the corner score if (x,y) is a FAST-9/16 corner at `threshold`, else 0.
static inline int fast_corner_score(const uint8_t* p, const int* off, int threshold) {
    const int v = p[0];
    // quick reject: any 9-arc of 16 contains >= 2 of the 4 compass pixels (0,4,8,12)
    int nb = 0, nd = 0;
    for (int k = 0; k < 16; k += 4) { int d = v - p[off[k]]; nb += d > threshold; nd += d < -threshold; }
    if (nb < 2 && nd < 2) return 0;
    int d[25];
    for (int k = 0; k < 16; ++k) d[k] = v - p[off[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int A = -256, B = -256;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { mn = d[k + j] < mn ? d[k + j] : mn; mx = d[k + j] > mx ? d[k + j] : mx; }
        if (mn > A) A = mn;     // max over arcs of min(v - p)
        if (-mx > B) B = -mx;   // max over arcs of min(p - v)
    }
    const int score = (A > B ? A : B) - 1;
    return score >= threshold ? score : 0;
}

struct FastKp { int x, y, score; };

static inline void fast9_16_nms(const Img& im, int threshold, std::vector<FastKp>& out) {
    out.clear();
    const int w = im.w, h = im.h;
    if (w < 7 || h < 7) return;
    int off[16];
    for (int k = 0; k < 16; ++k) off[k] = FAST_DY[k] * (int)im.stride + FAST_DX[k];
    std::vector<int> sc((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y) {
        const uint8_t* r = im.row(y);
        for (int x = 3; x < w - 3; ++x) sc[(size_t)y * w + x] = fast_corner_score(r + x, off, threshold);
    }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            const int s = sc[(size_t)y * w + x];
            if (s == 0 && threshold > 0) continue;
            if (s < threshold) continue;
            const int* c = &sc[(size_t)y * w + x];
            if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] && s > c[w - 1] && s > c[w] && s > c[w + 1])
                out.push_back({x, y, s});
        }
}

// First-party SWAR Hamming distance over 8 int32 words — include/Matcher.hpp:19-33.
static inline unsigned hamming256(const uint8_t* a, const uint8_t* b) {
    unsigned dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y; std::memcpy(&x, a + 4 * i, 4); std::memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24;
    }
    return dist;
}

}  // namespace ora
