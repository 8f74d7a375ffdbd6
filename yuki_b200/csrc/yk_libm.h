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

// ---- atanf / atan2f / acosf ------------------------------------------------------------------------
// glibc 2.39 still ships the fdlibm single-precision routines for these (sysdeps/ieee754/flt-32/
// s_atanf.c, e_atan2f.c, e_acosf.c; compiled without contraction): argument reduction to one of four
// intervals plus an odd/even split polynomial for atan, a rational approximation (with a split square
// root above 0.5) for acos. Only + - * / sqrt in f32, so the device evaluates them exactly when compiled
// --fmad=false -prec-div=true -prec-sqrt=true. Used by Sphere::intersect's (phi, theta)
// (shapes/sphere.rs:86-93); tests/test_libm.py compares with the host libm.
YK_HD float from_bits_(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float x;
    memcpy(&x, &u, 4);
    return x;
#endif
}
YK_HD float sqrt_(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return __builtin_sqrtf(x);
#endif
}
YK_HD float div_(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}

YK_HD float atanf_glibc(float x) {
    const float a0 = 3.3333334327e-01f, a1 = -2.0000000298e-01f, a2 = 1.4285714924e-01f, a3 = -1.1111110449e-01f,
                a4 = 9.0908870101e-02f, a5 = -7.6918758452e-02f, a6 = 6.6610731184e-02f, a7 = -5.8335702866e-02f,
                a8 = 4.9768779427e-02f, a9 = -3.6531571299e-02f, a10 = 1.6285819933e-02f;
    const uint32_t hx = bits_(x), ix = hx & 0x7fffffffu;
    const bool neg = (hx >> 31) != 0;
    float hi = 0.0f, lo = 0.0f;  // atan of the interval's centre (0.5, 1, 1.5, inf), split in two floats
    bool reduced = true;
    if (ix >= 0x4c000000u) {  // |x| >= 2^25
        if (ix > 0x7f800000u) return x + x;
        return neg ? -1.5707962513e+00f - 7.5497894159e-08f : 1.5707962513e+00f + 7.5497894159e-08f;
    }
    if (ix < 0x3ee00000u) {  // |x| < 7/16
        if (ix < 0x31000000u) return x;  // |x| < 2^-29
        reduced = false;
    } else {
        x = from_bits_(ix);
        if (ix < 0x3f980000u) {
            if (ix < 0x3f300000u) {
                hi = 4.6364760399e-01f; lo = 5.0121582440e-09f;
                x = div_(2.0f * x - 1.0f, 2.0f + x);
            } else {
                hi = 7.8539812565e-01f; lo = 3.7748947079e-08f;
                x = div_(x - 1.0f, x + 1.0f);
            }
        } else if (ix < 0x401c0000u) {
            hi = 9.8279368877e-01f; lo = 3.4473217170e-08f;
            x = div_(x - 1.5f, 1.0f + 1.5f * x);
        } else {
            hi = 1.5707962513e+00f; lo = 7.5497894159e-08f;
            x = div_(-1.0f, x);
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
    const float s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
    if (!reduced) return x - x * (s1 + s2);
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return neg ? -r : r;
}

YK_HD float atan2f_glibc(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
                pi_lo = -8.7422776573e-08f;
    const uint32_t hx = bits_(x), hy = bits_(y), ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    if (ix > 0x7f800000u || iy > 0x7f800000u) return x + y;
    if (hx == 0x3f800000u) return atanf_glibc(y);
    const uint32_t m = (hy >> 31) | ((hx >> 30) & 2u);  // 2*sign(x) + sign(y)
    if (iy == 0) {
        if (m < 2) return y;
        return m == 2 ? pi + tiny : -pi - tiny;
    }
    if (ix == 0) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000u) {
        if (iy == 0x7f800000u) {
            switch (m) {
                case 0: return pi_o_4 + tiny;
                case 1: return -pi_o_4 - tiny;
                case 2: return 3.0f * pi_o_4 + tiny;
                default: return -3.0f * pi_o_4 - tiny;
            }
        }
        switch (m) {
            case 0: return 0.0f;
            case 1: return -0.0f;
            case 2: return pi + tiny;
            default: return -pi - tiny;
        }
    }
    if (iy == 0x7f800000u) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int32_t k = ((int32_t)iy - (int32_t)ix) >> 23;
    float z;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if ((hx >> 31) && k < -60) z = 0.0f;
    else z = atanf_glibc(from_bits_(bits_(div_(y, x)) & 0x7fffffffu));
    switch (m) {
        case 0: return z;
        case 1: return from_bits_(bits_(z) ^ 0x80000000u);
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

YK_HD float acosf_glibc(float x) {
    const float pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f,
                pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
                pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
                qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    const uint32_t hx = bits_(x), ix = hx & 0x7fffffffu;
    const bool neg = (hx >> 31) != 0;
    if (ix == 0x3f800000u) return neg ? pi + 2.0f * pio2_lo : 0.0f;
    if (ix > 0x3f800000u) return div_(x - x, x - x);
    if (ix < 0x3f000000u) {  // |x| < 0.5
        if (ix <= 0x32800000u) return pio2_hi + pio2_lo;
        const float z = x * x;
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float r = div_(p, q);
        return pio2_hi - (x - (pio2_lo - r * x));
    }
    if (neg) {  // x < -0.5
        const float z = (1.0f + x) * 0.5f;
        const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        const float s = sqrt_(z);
        const float r = div_(p, q);
        const float w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    }
    const float z = (1.0f - x) * 0.5f;
    const float s = sqrt_(z);
    const float df = from_bits_(bits_(s) & 0xfffff000u);
    const float c = div_(z - df * df, s + df);
    const float p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
    const float q = 1.0f + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
    const float r = div_(p, q);
    const float w = r * s + c;
    return 2.0f * (df + w);
}

// ---- logf -------------------------------------------------------------------------------------------
// glibc 2.39's logf is the Arm Optimized Routines one: k = exponent, a 16-entry table of (1/c, log c) by the top four
// mantissa bits, r = z/c - 1 and a cubic in r, all in f64, rounded once. Used by roughness_to_alpha
// (trowbridge_reitz.rs:22-30) when a Metal / Glossy roughness comes from an image texture (constant textures are
// converted on the host). tests/test_libm.py compares every positive float with the host libm.
#if defined(__CUDACC__)  // the table lives in device memory there (a local copy would cost every caller 256 B of stack)
#define YK_LOGF_FN __device__ inline
#define YK_LOGF_TABLE static __device__ const
#else
#define YK_LOGF_FN inline
#define YK_LOGF_TABLE static const
#endif
YK_LOGF_TABLE double kLogfTable[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
    {0x1.49539f0f010bp+0, -0x1.01eae7f513a67p-2},  {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
    {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8eap+0, -0x1.1aa2bc79c81p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
    {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1p+0, 0x0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5},  {0x1.ca4b31f026aap-1, 0x1.c5e53aa362eb4p-4},
    {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d22477p-3},
    {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2},  {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2},
};
YK_LOGF_FN float logf_glibc(float x) {
    const double ln2 = 0x1.62e42fefa39efp-1, a0 = -0x1.00ea348b88334p-2, a1 = 0x1.5575b0be00b6ap-2, a2 = -0x1.ffffef20a4123p-2;
    uint32_t ix = bits_(x);
    if (ix == 0x3f800000u) return 0.0f;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {  // zero, subnormal, negative, inf, NaN
        if (ix * 2u == 0u) return -from_bits_(0x7f800000u);
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2u >= 0xff000000u) return from_bits_(0x7fc00000u);
        ix = bits_(x * 0x1p23f) - (23u << 23);
    }
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (int)((tmp >> 19) & 15u);
    const int k = (int32_t)tmp >> 23;
    const double z = (double)from_bits_(ix - (tmp & 0xff800000u));
    const double r = mul_(z, kLogfTable[i][0]) - 1.0;
    const double y0 = kLogfTable[i][1] + mul_((double)k, ln2);
    const double r2 = mul_(r, r);
    double y = mul_(a1, r) + a2;
    y = mul_(a0, r2) + y;
    y = mul_(y, r2) + (y0 + r);
    return (float)y;
}

}  // namespace yklibm
