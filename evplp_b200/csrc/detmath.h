// detmath.h -- the deterministic "libm" of the EVPLP hot path.
//
// Why it exists: emitted photon/VPL counts depend on Russian-roulette
// decisions, which depend on flux, which depends on powf/sinf/cosf results
// (reference: realtimetechniques/rtmaterial.cuh:120-155, lighttracing.cu:164-166).
// CUDA libdevice and glibc differ by ulps, so the sm_100a kernels and the CPU oracle
// both call THESE functions instead.  Every function below is built only from IEEE-754
// correctly-rounded primitives (+ - * / sqrt, explicit fma, rint, int<->float casts),
// so the result is bit-identical on x86-64 and on sm_100a provided neither compiler
// contracts a*b+c on its own (nvcc -fmad=false, g++ -ffp-contract=off).
//
// This header plays the role of libm for both sides; it contains no rendering logic.
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define EVPLP_HD __host__ __device__ __forceinline__
#define EVPLP_HD_NOINLINE __host__ __device__
#else
#define EVPLP_HD inline
#define EVPLP_HD_NOINLINE
#endif

namespace evplp {

constexpr float kPi = 3.14159265358979323846f;     // OptiX M_PIf
constexpr float kInvPi = 0.318309886183790671537767526745028724068919291480912897495f;  // rtmath.cuh:12

EVPLP_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
EVPLP_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
EVPLP_HD uint64_t d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
EVPLP_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}

// IEEE correctly-rounded primitives under one name per side.
EVPLP_HD float det_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}
EVPLP_HD double det_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
EVPLP_HD float det_sqrtf(float x) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return __builtin_sqrtf(x);
#endif
}
EVPLP_HD float det_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}
// fmaxf/fminf with the CUDA/C99 NaN rule (return the non-NaN operand).
EVPLP_HD float det_max(float a, float b) { return fmaxf(a, b); }
EVPLP_HD float det_min(float a, float b) { return fminf(a, b); }

// ---------------------------------------------------------------------------------
// sinf / cosf for moderate arguments (|x| <= ~1e3; the path uses [0, 2*pi]).
// Cody-Waite 3-term reduction by pi/2 with fma, then degree-7 / degree-8 minimax
// polynomials on [-pi/4, pi/4].
// ---------------------------------------------------------------------------------
EVPLP_HD void det_sincosf(float x, float* s, float* c) {
    const float kTwoOverPi = 0x1.45f306p-1f;
    const float kPio2Hi = 0x1.921fb6p+0f;
    const float kPio2Mid = -0x1.777a5cp-25f;
    const float kPio2Lo = -0x1.ee59dap-50f;
    float k = rintf(x * kTwoOverPi);
    float r = det_fma(-k, kPio2Hi, x);
    r = det_fma(-k, kPio2Mid, r);
    r = det_fma(-k, kPio2Lo, r);
    int q = (int)k;
    float r2 = r * r;
    // sin(r) = r + r^3 * (S1 + r^2 (S2 + r^2 S3))
    float ps = det_fma(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = det_fma(r2, ps, -1.6666654611e-1f);
    float sr = det_fma(r * r2, ps, r);
    // cos(r) = 1 - r^2/2 + r^4 (C1 + r^2 (C2 + r^2 C3))
    float pc = det_fma(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = det_fma(r2, pc, 4.166664568298827e-2f);
    float cr = det_fma(r2 * r2, pc, det_fma(r2, -0.5f, 1.0f));
    float ss = (q & 1) ? cr : sr;
    float cc = (q & 1) ? sr : cr;
    if (q & 2) ss = -ss;
    if ((q + 1) & 2) cc = -cc;
    *s = ss;
    *c = cc;
}
EVPLP_HD float det_sinf(float x) { float s, c; det_sincosf(x, &s, &c); return s; }
EVPLP_HD float det_cosf(float x) { float s, c; det_sincosf(x, &s, &c); return c; }

// ---------------------------------------------------------------------------------
// powf(x, y) for x > 0: exp2(y * log2(x)) evaluated in double so that the float result
// is within 1 ulp for every exponent the path uses (Phong exponents up to thousands).
// B200 runs FP64 at half the FP32 rate, so ~30 double ops are comparable to libdevice's
// float-float powf.
// ---------------------------------------------------------------------------------
EVPLP_HD double det_log2_pos(double x) {
    // x is a positive normal double (a widened float, possibly a widened subnormal float,
    // which is a normal double).
    uint64_t ux = d2u(x);
    int e = (int)((ux >> 52) & 0x7ff) - 1023;
    uint64_t mant = ux & 0x000fffffffffffffULL;
    // m in [1,2); fold to [sqrt(1/2), sqrt(2))
    if (mant > 0x6a09e667f3bcdULL) {  // m > sqrt(2)
        e += 1;
        ux = mant | 0x3fe0000000000000ULL;  // m/2 in [0.5,1)
    } else {
        ux = mant | 0x3ff0000000000000ULL;
    }
    double m = u2d(ux);
    double f = m - 1.0;
    double s = f / (2.0 + f);
    double z = s * s;
    // ln(m) = 2 atanh(s) = 2 s (1 + z/3 + z^2/5 + ... + z^7/15), |z| <= 0.0295
    double p = 1.0 / 15.0;
    p = det_fma(p, z, 1.0 / 13.0);
    p = det_fma(p, z, 1.0 / 11.0);
    p = det_fma(p, z, 1.0 / 9.0);
    p = det_fma(p, z, 1.0 / 7.0);
    p = det_fma(p, z, 1.0 / 5.0);
    p = det_fma(p, z, 1.0 / 3.0);
    p = det_fma(p, z, 1.0);
    double lnm = 2.0 * s * p;
    const double kInvLn2 = 0x1.71547652b82fep+0;
    return det_fma(lnm, kInvLn2, (double)e);
}

EVPLP_HD double det_exp2_d(double t) {
    // caller clamps t to [-1100, 1100]
    double k = rint(t);
    double f = t - k;  // [-0.5, 0.5]
    const double kLn2 = 0x1.62e42fefa39efp-1;
    double u = f * kLn2;
    double p = 0x1.1eed8eff8d898p-29;  // 1/12!
    p = det_fma(p, u, 0x1.ae64567f544e4p-26);
    p = det_fma(p, u, 0x1.27e4fb7789f5cp-22);
    p = det_fma(p, u, 0x1.71de3a556c734p-19);
    p = det_fma(p, u, 0x1.a01a01a01a01ap-16);
    p = det_fma(p, u, 0x1.a01a01a01a01ap-13);
    p = det_fma(p, u, 0x1.6c16c16c16c17p-10);
    p = det_fma(p, u, 0x1.1111111111111p-7);
    p = det_fma(p, u, 0x1.5555555555555p-5);
    p = det_fma(p, u, 0x1.5555555555555p-3);
    p = det_fma(p, u, 0.5);
    p = det_fma(p, u, 1.0);
    p = det_fma(p, u, 1.0);
    // scale by 2^k, k in [-1100, 1100]: split to stay inside the double exponent range
    int ki = (int)k;
    int k1 = ki / 2, k2 = ki - k1;
    double s1 = u2d((uint64_t)(k1 + 1023) << 52);
    double s2 = u2d((uint64_t)(k2 + 1023) << 52);
    return p * s1 * s2;
}

EVPLP_HD float det_powf(float x, float y) {
    // Domain used by the path: x > 0 (cosines above 1e-6, uniforms in (0,1]), y finite.
    if (y == 0.0f) return 1.0f;
    if (x == 1.0f) return 1.0f;
    if (!(x > 0.0f)) return (x == 0.0f) ? (y > 0.0f ? 0.0f : INFINITY) : NAN;
    double t = (double)y * det_log2_pos((double)x);
    if (t > 1100.0) t = 1100.0;
    if (t < -1100.0) t = -1100.0;
    return (float)det_exp2_d(t);
}

// ---------------------------------------------------------------------------------
// asinf(x) for x in [0, 1): Taylor series in double on [0, 0.5], the usual
// pi/2 - 2 asin(sqrt((1-x)/2)) reflection above.
// ---------------------------------------------------------------------------------
EVPLP_HD double det_asin_small(double x) {  // |x| <= 0.5
    double z = x * x;
    double p = 0x1.fcaf8fb6db6dbp-9;
    p = det_fma(p, z, 0x1.15ee9d45d1746p-8);
    p = det_fma(p, z, 0x1.31683bdef7bdfp-8);
    p = det_fma(p, z, 0x1.51ba308d3dcb1p-8);
    p = det_fma(p, z, 0x1.782dda12f684cp-8);
    p = det_fma(p, z, 0x1.a6863d70a3d71p-8);
    p = det_fma(p, z, 0x1.df3bd37a6f4dfp-8);
    p = det_fma(p, z, 0x1.12ef3cf3cf3cfp-7);
    p = det_fma(p, z, 0x1.3fde50d79435ep-7);
    p = det_fma(p, z, 0x1.7a87878787878p-7);
    p = det_fma(p, z, 0x1.c99999999999ap-7);
    p = det_fma(p, z, 0x1.1c4ec4ec4ec4fp-6);
    p = det_fma(p, z, 0x1.6e8ba2e8ba2e9p-6);
    p = det_fma(p, z, 0x1.f1c71c71c71c7p-6);
    p = det_fma(p, z, 0x1.6db6db6db6db7p-5);
    p = det_fma(p, z, 0x1.3333333333333p-4);
    p = det_fma(p, z, 0x1.5555555555555p-3);
    return det_fma(x * z, p, x);
}
EVPLP_HD double det_sqrt_d(double x) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return __builtin_sqrt(x);
#endif
}
EVPLP_HD float det_asinf(float xf) {
    double x = (double)xf;
    double ax = x < 0.0 ? -x : x;
    double r;
    if (ax <= 0.5) {
        r = det_asin_small(ax);
    } else {
        if (ax > 1.0) return NAN;
        const double kPio2 = 0x1.921fb54442d18p+0;
        double h = det_sqrt_d((1.0 - ax) * 0.5);
        r = kPio2 - 2.0 * det_asin_small(h);
    }
    return (float)(x < 0.0 ? -r : r);
}

}  // namespace evplp
