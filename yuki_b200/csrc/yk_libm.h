// sinf / cosf with the exact results of glibc 2.39's x86-64 FMA variants (the libm that Rust's f32::sin /
// f32::cos resolve to on the Linux hosts yuki runs on; yuki/src/sampling/mod.rs:86, bsdfs/mod.rs:280).
//
// glibc's implementation is the Arm Optimized Routines single-precision sincos (published, MIT): the
// argument is widened to f64, reduced by round(x * 2/pi) for |x| < 120, and a degree-7 / degree-8
// polynomial pair is evaluated in f64 with fused multiply-adds, then rounded once to f32. The operation
// order below (including which products are fused) restates that algorithm as glibc 2.39 compiles it, so
// device results are bit-identical to the host libm's; tests/test_libm.py checks 2^24+ arguments.
// |x| >= 120 (never produced by the sampling code: theta in [-pi/4, 3pi/4], phi in [0, 2pi)) falls back to
// f64 sin/cos rounded to f32.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define YK_HD __host__ __device__ inline
#else
#define YK_HD inline
#endif

namespace yklibm {

YK_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
YK_HD double mul_(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
YK_HD uint32_t bits_(float x) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u;
    memcpy(&u, &x, 4);
    return u;
#endif
}

struct Poly {
    double c0, c1, s1, c2, s2, c3, s3, c4;
};
// Quadrants 0/1 use +cos, quadrants 2/3 use -cos (sign folded into the coefficients); the sine
// coefficients are shared and the sine's sign comes from sign[n & 3].
YK_HD Poly poly_for(int negate_cos) {
    const double sgn = negate_cos ? -1.0 : 1.0;
    return Poly{sgn * 0x1p0,
                sgn * -0x1.ffffffd0c621cp-2,
                -0x1.555545995a603p-3,
                sgn * 0x1.55553e1068f19p-5,
                0x1.1107605230bc4p-7,
                sgn * -0x1.6c087e89a359dp-10,
                -0x1.994eb3774cf24p-13,
                sgn * 0x1.99343027bf8c3p-16};
}
YK_HD float sin_poly(double x, double x2, const Poly& p) {
    const double x3 = mul_(x, x2);
    const double s1 = fma_(x2, p.s3, p.s2);
    const double x7 = mul_(x3, x2);
    const double s = fma_(x3, p.s1, x);
    return (float)fma_(s1, x7, s);
}
YK_HD float cos_poly(double x2, const Poly& p) {
    const double x4 = mul_(x2, x2);
    const double a = fma_(x2, p.c1, p.c0);
    const double c2 = fma_(x2, p.c4, p.c3);
    const double x6 = mul_(x4, x2);
    const double c = fma_(x4, p.c2, a);
    return (float)fma_(c2, x6, c);
}
// r = x - n * pi/2 with n = round(x * 2/pi), both in f64 (reduce_fast)
YK_HD double reduce_fast(double x, int* n_out) {
    const double r = mul_(x, 0x1.45F306DC9C883p+23);
    const int n = ((int32_t)r + 0x800000) >> 24;
    *n_out = n;
    return fma_(-(double)n, 0x1.921FB54442D18p0, x);
}
YK_HD double quadrant_sign(int n) { return ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0; }

YK_HD float sinf_glibc(float y) {
    const uint32_t top = (bits_(y) >> 20) & 0x7ffu;
    double x = (double)y;
    if (top <= 0x3f3u) {  // |y| < pi/4
        if (top <= 0x397u) return y;  // |y| < 2^-12
        return sin_poly(x, mul_(x, x), poly_for(0));
    }
    if (top <= 0x42eu) {  // |y| < 120
        int n;
        x = reduce_fast(x, &n);
        const Poly p = poly_for(n & 2);
        const double x2 = mul_(x, x);
        if (n & 1) return cos_poly(x2, p);
        return sin_poly(mul_(x, quadrant_sign(n)), x2, p);
    }
    return (float)sin((double)y);
}
YK_HD float cosf_glibc(float y) {
    const uint32_t top = (bits_(y) >> 20) & 0x7ffu;
    double x = (double)y;
    if (top <= 0x3f3u) {
        if (top <= 0x397u) return 1.0f;
        return cos_poly(mul_(x, x), poly_for(0));
    }
    if (top <= 0x42eu) {
        int n;
        x = reduce_fast(x, &n);
        const Poly p = poly_for(n & 2);
        const double x2 = mul_(x, x);
        if (n & 1) return sin_poly(mul_(x, quadrant_sign(n)), x2, p);
        return cos_poly(x2, p);
    }
    return (float)cos((double)y);
}

}  // namespace yklibm
