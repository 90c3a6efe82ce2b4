// Device float routines that must reproduce the reference's x86-64 results bit for bit. Every operation is an explicit
// round-to-nearest intrinsic so that nvcc never contracts a*b+c into an FMA (the reference is built without -march,
// CMakeLists.txt:18-19, so it has separate mul/add roundings).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mcv {

// cvRound(float) == cvtss2si == round-half-even.
__device__ __forceinline__ int cv_round_f(float v) { return __float2int_rn(v); }

// cv::fastAtan2(y, x) — called by IC_Angle, ORBextractor.cc:97. OpenCV core/mathfuncs_core.simd.hpp atan_f32:
// odd 7th-order polynomial in degrees; returns [0, 360).
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float scale = 57.295779513082321f;  // (float)(180/CV_PI)
    const float p1 = __fmul_rn(0.9997878412794807f, scale), p3 = __fmul_rn(-0.3258083974640975f, scale),
                p5 = __fmul_rn(0.1555786518463281f, scale), p7 = __fmul_rn(-0.04432655554792128f, scale);
    const float eps = 2.2204460492503131e-16f;  // (float)DBL_EPSILON
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

// glibc 2.39 sincosf (sysdeps/ieee754/flt-32/s_sincosf.c, the ARM optimized-routines algorithm) for |y| < 120:
// `(float)cos(angle)` / `(float)sin(angle)` at ORBextractor.cc:102-103 take a float argument under `using namespace std`,
// so they resolve to the float overloads and GCC fuses them into one sincosf call. Double-precision range reduction by
// pi/2 and two minimax polynomials, rounded once to float. This port was checked against libm over every float in
// [0, 6.5] (1 087 373 313 values, 0 mismatches; with and without FMA contraction) — DESIGN.md "sincosf".
__device__ __forceinline__ void sincosf_glibc(float y, float* sinp, float* cosp) {
    const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;
    const double C0 = 1.0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                 C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ffu;
    double x = (double)y;
    if (top < ((__float_as_uint(0x1.921FB6p-1f) >> 20) & 0x7ffu)) {  // |y| < pi/4
        if (top < ((__float_as_uint(0x1p-12f) >> 20) & 0x7ffu)) { *sinp = y; *cosp = 1.0f; return; }
        double x2 = __dmul_rn(x, x);
        double x4 = __dmul_rn(x2, x2), x3 = __dmul_rn(x2, x);
        double c2 = __dadd_rn(C3, __dmul_rn(x2, C4)), s1 = __dadd_rn(S2, __dmul_rn(x2, S3));
        double c1 = __dadd_rn(C0, __dmul_rn(x2, C1));
        double x5 = __dmul_rn(x3, x2), x6 = __dmul_rn(x4, x2);
        double s = __dadd_rn(x, __dmul_rn(x3, S1)), c = __dadd_rn(c1, __dmul_rn(x4, C2));
        *sinp = (float)__dadd_rn(s, __dmul_rn(x5, s1));
        *cosp = (float)__dadd_rn(c, __dmul_rn(x6, c2));
        return;
    }
    // reduce_fast: quadrant n, x in [-pi/4, pi/4]
    double r = __dmul_rn(x, HPI_INV);
    int n = ((int32_t)r + 0x800000) >> 24;
    x = __dsub_rn(x, __dmul_rn((double)n, HPI));
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const double f = (n & 2) ? -1.0 : 1.0;  // second table = cosine coefficients negated
    double xs = __dmul_rn(x, sgn), x2 = __dmul_rn(x, x);
    double x4 = __dmul_rn(x2, x2), x3 = __dmul_rn(x2, xs);
    double c2 = __dadd_rn(f * C3, __dmul_rn(x2, f * C4)), s1 = __dadd_rn(S2, __dmul_rn(x2, S3));
    double c1 = __dadd_rn(f * C0, __dmul_rn(x2, f * C1));
    double x5 = __dmul_rn(x3, x2), x6 = __dmul_rn(x4, x2);
    double s = __dadd_rn(xs, __dmul_rn(x3, S1)), c = __dadd_rn(c1, __dmul_rn(x4, f * C2));
    float sv = (float)__dadd_rn(s, __dmul_rn(x5, s1));
    float cv = (float)__dadd_rn(c, __dmul_rn(x6, c2));
    if (n & 1) { *sinp = cv; *cosp = sv; } else { *sinp = sv; *cosp = cv; }
}

__device__ __forceinline__ int hamming256(const uint4& a0, const uint4& a1, const uint4& b0, const uint4& b1) {
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

}  // namespace mcv
