// CPU restatement (TEST INFRASTRUCTURE — only tests/, smoke() and bench.py's cpu_baseline may use it) of the optical-flow
// pre-seeding of the Frame constructor: KL_Track (src/Frame.cpp:34-76) calls
//     cv::calcOpticalFlowPyrLK(obj1->img, obj2->img, pts, next_pts, res, err, Size(10, 10), 1,
//                              TermCriteria(COUNT + EPS, 10, 0.01), 0, 0.001)                       (src/Frame.cpp:52-54)
// and keeps a point when res[i] > 0 && err[i] < 1 (src/Frame.cpp:57-58).
//
// OpenCV is not vendored in the reference; this restates the published algorithm of OpenCV 4's video/src/lkpyramid.cpp
// (buildOpticalFlowPyramid, calcScharrDeriv, LKTrackerInvoker) and imgproc's pyrDown for 8-bit single-channel images.
// Integer parts (pyramid, Scharr derivatives, 14-bit bilinear window interpolation) are exact by construction. The float sums
// (A11, A12, A22, b1, b2) follow the SSE2 path that an x86-64 OpenCV build (baseline SSE3, which is what the reference links
// on its x86 host and what the cv2 wheel is) executes: four lanes own pixels x and x + 4 of every window row (b1 / b2: the two
// products are added as int32 by v_dotprod before the conversion to float), pixels 8 and 9 run through the scalar tail, and
// v_reduce_sum adds (q0 + q2) + (q1 + q3). PINNED BIT-EXACTLY against cv2 4.13: tests/golden/make_lk_golden.py ->
// tests/golden/lk_golden.npz, tests/test_oracle_golden.py::test_lk_oracle_against_cv2_fixture (status, positions and err
// identical on 4 scene pairs x 1500 points, points outside the image included). One deliberate definition: a window whose
// last row is the bottom ring row reads one row past OpenCV's buffer (an over-read in OpenCV itself); here that row is zeros.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace ora_lk {

static inline int reflect101(int p, int len) {          // cv::borderInterpolate(p, len, BORDER_REFLECT_101)
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * (len - 1) - p; }
    return p;
}

// cv::pyrDown, CV_8UC1, BORDER_REFLECT_101: separable [1 4 6 4 1], (sum + 128) >> 8; dst = ((w + 1) / 2, (h + 1) / 2)
static void pyr_down(const std::vector<uint8_t>& src, int w, int h, std::vector<uint8_t>& dst, int& dw, int& dh) {
    dw = (w + 1) / 2; dh = (h + 1) / 2;
    dst.assign((size_t)dw * dh, 0);
    std::vector<int> rows((size_t)5 * dw);
    for (int y = 0; y < dh; ++y) {
        for (int k = 0; k < 5; ++k) {
            const uint8_t* s = &src[(size_t)reflect101(2 * y - 2 + k, h) * w];
            int* r = &rows[(size_t)k * dw];
            for (int x = 0; x < dw; ++x)
                r[x] = s[reflect101(2 * x, w)] * 6 + (s[reflect101(2 * x - 1, w)] + s[reflect101(2 * x + 1, w)]) * 4 + s[reflect101(2 * x - 2, w)] +
                       s[reflect101(2 * x + 2, w)];
        }
        for (int x = 0; x < dw; ++x)
            dst[(size_t)y * dw + x] = (uint8_t)((rows[x] + rows[4 * dw + x] + (rows[dw + x] + rows[3 * dw + x]) * 4 + rows[2 * dw + x] * 6 + 128) >> 8);
    }
}

constexpr int WIN = 10;        // winSize (both), also the border of every level buffer
constexpr int W_BITS = 14;

struct Level {
    int w = 0, h = 0, stride = 0;          // stride of both buffers in elements; origin of the level at (WIN, WIN)
    std::vector<uint8_t> img;              // (h + 2 WIN) x stride, BORDER_REFLECT_101 ring (copyMakeBorder of buildOpticalFlowPyramid)
    std::vector<int16_t> deriv;            // (h + 2 WIN) x stride x 2 (dx, dy), zero ring (BORDER_CONSTANT)
    const uint8_t* I(int x, int y) const { return &img[(size_t)(y + WIN) * stride + x + WIN]; }
    const int16_t* D(int x, int y) const { return &deriv[((size_t)(y + WIN) * stride + x + WIN) * 2]; }
};

static void make_level(const std::vector<uint8_t>& src, int w, int h, bool with_deriv, Level& L) {
    L.w = w; L.h = h; L.stride = w + 2 * WIN;
    L.img.assign((size_t)(h + 2 * WIN + 1) * L.stride, 0);   // + one zero row: the window's y + 1 taps of a point at the very bottom
    for (int y = -WIN; y < h + WIN; ++y)
        for (int x = -WIN; x < w + WIN; ++x) L.img[(size_t)(y + WIN) * L.stride + x + WIN] = src[(size_t)reflect101(y, h) * w + reflect101(x, w)];
    if (!with_deriv) return;
    // calcScharrDeriv: smoothing (3, 10, 3) x difference (-1, 0, 1), image borders by reflection (lkpyramid.cpp)
    L.deriv.assign((size_t)(h + 2 * WIN + 1) * L.stride * 2, 0);
    std::vector<int> t0(w + 2), t1(w + 2);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s0 = &src[(size_t)(y > 0 ? y - 1 : h > 1 ? 1 : 0) * w];
        const uint8_t* s1 = &src[(size_t)y * w];
        const uint8_t* s2 = &src[(size_t)(y < h - 1 ? y + 1 : h > 1 ? h - 2 : 0) * w];
        for (int x = 0; x < w; ++x) { t0[x + 1] = (s0[x] + s2[x]) * 3 + s1[x] * 10; t1[x + 1] = s2[x] - s0[x]; }
        const int x0 = w > 1 ? 1 : 0, x1 = w > 1 ? w - 2 : 0;
        t0[0] = t0[x0 + 1]; t0[w + 1] = t0[x1 + 1]; t1[0] = t1[x0 + 1]; t1[w + 1] = t1[x1 + 1];
        int16_t* d = &L.deriv[((size_t)(y + WIN) * L.stride + WIN) * 2];
        for (int x = 0; x < w; ++x) {
            d[2 * x] = (int16_t)(t0[x + 2] - t0[x]);
            d[2 * x + 1] = (int16_t)((t1[x + 2] + t1[x]) * 3 + t1[x + 1] * 10);
        }
    }
}

static inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
static inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// LKTrackerInvoker::operator() for one point on one level.
static void track_level(const Level& I, const Level& J, int level, int max_level, float px, float py, float& nx, float& ny, uint8_t& status,
                        float& err, int max_count, double eps2, float min_eig) {
    const float halfx = (WIN - 1) * 0.5f, halfy = (WIN - 1) * 0.5f;
    const float lvl_scale = (float)(1. / (1 << level));
    float prevx = px * lvl_scale, prevy = py * lvl_scale;
    float nextx, nexty;
    if (level == max_level) { nextx = prevx; nexty = prevy; }
    else { nextx = nx * 2.f; nexty = ny * 2.f; }
    nx = nextx; ny = nexty;
    prevx -= halfx; prevy -= halfy;
    const int ipx = cv_floor_f(prevx), ipy = cv_floor_f(prevy);
    if (ipx < -WIN || ipx >= I.w || ipy < -WIN || ipy >= I.h) {
        if (level == 0) { status = 0; err = 0; }
        return;
    }
    float a = prevx - ipx, b = prevy - ipy;
    const float FLT_SCALE = 1.f / (1 << 20);
    int iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS));
    int iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
    int iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS));
    int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
    int16_t Iwin[WIN * WIN], dIwin[WIN * WIN * 2];
    float iA11 = 0, iA12 = 0, iA22 = 0;
    float qA11[4] = {0, 0, 0, 0}, qA12[4] = {0, 0, 0, 0}, qA22[4] = {0, 0, 0, 0};   // SSE lanes: pixels x and x + 4 of every row, x < 4
    for (int y = 0; y < WIN; ++y) {
        const uint8_t* src = I.I(ipx, ipy + y);
        const uint8_t* src1 = I.I(ipx, ipy + y + 1);
        const int16_t* ds = I.D(ipx, ipy + y);
        const int16_t* ds1 = I.D(ipx, ipy + y + 1);
        for (int x = 0; x < WIN; ++x) {
            const int ival = descale(src[x] * iw00 + src[x + 1] * iw01 + src1[x] * iw10 + src1[x + 1] * iw11, W_BITS - 5);
            const int ixval = descale(ds[2 * x] * iw00 + ds[2 * x + 2] * iw01 + ds1[2 * x] * iw10 + ds1[2 * x + 2] * iw11, W_BITS);
            const int iyval = descale(ds[2 * x + 1] * iw00 + ds[2 * x + 3] * iw01 + ds1[2 * x + 1] * iw10 + ds1[2 * x + 3] * iw11, W_BITS);
            Iwin[y * WIN + x] = (int16_t)ival;
            dIwin[(y * WIN + x) * 2] = (int16_t)ixval;
            dIwin[(y * WIN + x) * 2 + 1] = (int16_t)iyval;
            if (x < 8) {
                qA11[x & 3] += (float)(ixval * ixval);
                qA12[x & 3] += (float)(ixval * iyval);
                qA22[x & 3] += (float)(iyval * iyval);
            } else {
                iA11 += (float)(ixval * ixval);
                iA12 += (float)(ixval * iyval);
                iA22 += (float)(iyval * iyval);
            }
        }
    }
    iA11 += (qA11[0] + qA11[2]) + (qA11[1] + qA11[3]);      // v_reduce_sum
    iA12 += (qA12[0] + qA12[2]) + (qA12[1] + qA12[3]);
    iA22 += (qA22[0] + qA22[2]) + (qA22[1] + qA22[3]);
    const float A11 = iA11 * FLT_SCALE, A12 = iA12 * FLT_SCALE, A22 = iA22 * FLT_SCALE;
    float D = A11 * A22 - A12 * A12;
    const float minEig = (A22 + A11 - std::sqrt((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * WIN * WIN);
    if (minEig < min_eig || D < 1.1920928955078125e-07f) {
        if (level == 0) status = 0;
        return;
    }
    D = 1.f / D;
    nextx -= halfx; nexty -= halfy;
    float pdx = 0, pdy = 0;
    for (int j = 0; j < max_count; ++j) {
        const int inx = cv_floor_f(nextx), iny = cv_floor_f(nexty);
        if (inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h) {
            if (level == 0) status = 0;
            break;
        }
        a = nextx - inx; b = nexty - iny;
        iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS));
        iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
        iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS));
        iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
        float ib1 = 0, ib2 = 0;
        // SSE lanes: v_dotprod adds the products of pixels x and x + 4 (x < 4) as int32 before the conversion to float
        float qb1[4] = {0, 0, 0, 0}, qb2[4] = {0, 0, 0, 0};
        for (int y = 0; y < WIN; ++y) {
            const uint8_t* Jp = J.I(inx, iny + y);
            const uint8_t* Jp1 = J.I(inx, iny + y + 1);
            int diff[WIN];
            for (int x = 0; x < WIN; ++x)
                diff[x] = descale(Jp[x] * iw00 + Jp[x + 1] * iw01 + Jp1[x] * iw10 + Jp1[x + 1] * iw11, W_BITS - 5) - Iwin[y * WIN + x];
            for (int x = 0; x < 4; ++x) {
                qb1[x] += (float)(diff[x] * dIwin[(y * WIN + x) * 2] + diff[x + 4] * dIwin[(y * WIN + x + 4) * 2]);
                qb2[x] += (float)(diff[x] * dIwin[(y * WIN + x) * 2 + 1] + diff[x + 4] * dIwin[(y * WIN + x + 4) * 2 + 1]);
            }
            for (int x = 8; x < WIN; ++x) {
                ib1 += (float)(diff[x] * dIwin[(y * WIN + x) * 2]);
                ib2 += (float)(diff[x] * dIwin[(y * WIN + x) * 2 + 1]);
            }
        }
        ib1 += ((qb1[0] + qb1[2]) + 0.f) + ((qb1[1] + qb1[3]) + 0.f);   // qb0 + qb1, interleave, v_reduce_sum with a zero upper half
        ib2 += ((qb2[0] + qb2[2]) + 0.f) + ((qb2[1] + qb2[3]) + 0.f);
        const float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
        const float dx = (float)((A12 * b2 - A22 * b1) * D), dy = (float)((A12 * b1 - A11 * b2) * D);
        nextx += dx; nexty += dy;
        nx = nextx + halfx; ny = nexty + halfy;
        if ((double)dx * dx + (double)dy * dy <= eps2) break;
        if (j > 0 && std::abs(dx + pdx) < 0.01 && std::abs(dy + pdy) < 0.01) {
            nx -= dx * 0.5f; ny -= dy * 0.5f;
            break;
        }
        pdx = dx; pdy = dy;
    }
    if (status && level == 0) {
        const float qx = nx - halfx, qy = ny - halfy;
        const int inx = cv_floor_f(qx), iny = cv_floor_f(qy);
        if (inx < -WIN || inx >= J.w || iny < -WIN || iny >= J.h) { status = 0; return; }
        const float aa = qx - inx, bb = qy - iny;
        iw00 = (int)lrintf((1.f - aa) * (1.f - bb) * (1 << W_BITS));
        iw01 = (int)lrintf(aa * (1.f - bb) * (1 << W_BITS));
        iw10 = (int)lrintf((1.f - aa) * bb * (1 << W_BITS));
        iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
        float errval = 0.f;
        for (int y = 0; y < WIN; ++y) {
            const uint8_t* Jp = J.I(inx, iny + y);
            const uint8_t* Jp1 = J.I(inx, iny + y + 1);
            for (int x = 0; x < WIN; ++x) {
                const int diff = descale(Jp[x] * iw00 + Jp[x + 1] * iw01 + Jp1[x] * iw10 + Jp1[x + 1] * iw11, W_BITS - 5) - Iwin[y * WIN + x];
                errval += std::abs((float)diff);
            }
        }
        err = errval * 1.f / (32 * WIN * WIN);
    }
}

// cv::calcOpticalFlowPyrLK(prev, next, pts, ..., Size(10,10), maxLevel = 1, (COUNT+EPS, 10, 0.01), flags = 0, minEig = 0.001)
static void calc_lk(const uint8_t* prev, const uint8_t* next, int w, int h, int stride, const float* pts, int n, float* next_pts, uint8_t* status,
                    float* err) {
    const int max_count = 10;
    const double eps = 0.01, eps2 = eps * eps;
    const float min_eig = (float)0.001;
    std::vector<uint8_t> p0((size_t)w * h), n0((size_t)w * h), p1, n1;
    for (int y = 0; y < h; ++y) { memcpy(&p0[(size_t)y * w], prev + (size_t)y * stride, w); memcpy(&n0[(size_t)y * w], next + (size_t)y * stride, w); }
    int max_level = 1, w1 = 0, h1 = 0;
    pyr_down(p0, w, h, p1, w1, h1);
    pyr_down(n0, w, h, n1, w1, h1);
    if (w1 <= WIN || h1 <= WIN) max_level = 0;      // buildOpticalFlowPyramid stops before a level not larger than the window
    Level I[2], J[2];
    make_level(p0, w, h, true, I[0]); make_level(n0, w, h, false, J[0]);
    if (max_level == 1) { make_level(p1, w1, h1, true, I[1]); make_level(n1, w1, h1, false, J[1]); }
    for (int i = 0; i < n; ++i) { status[i] = 1; err[i] = 0; next_pts[2 * i] = 0; next_pts[2 * i + 1] = 0; }
    for (int level = max_level; level >= 0; --level)
        for (int i = 0; i < n; ++i)
            track_level(I[level], J[level], level, max_level, pts[2 * i], pts[2 * i + 1], next_pts[2 * i], next_pts[2 * i + 1], status[i], err[i],
                        max_count, eps2, min_eig);
}

}  // namespace ora_lk
