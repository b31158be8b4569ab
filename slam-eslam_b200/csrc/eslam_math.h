// eslam_math.h -- host/device arithmetic shared by the CUDA kernels and by the CPU oracle's
// "shared" arithmetic policy.
//
// Every function here is built ONLY from IEEE-754 correctly rounded primitives
// (+ - * / sqrt fma, int<->fp conversions, bit moves), with every contraction written
// out as an explicit fma().  Compile device code with `-fmad=false` and host code with
// `-ffp-contract=off` and the two sides produce bit-identical doubles.  That is what lets
// the GPU path be compared bit-for-bit with the oracle instead of "within libm ulps".
//
// Nothing in this file is taken from the reference; the reference leans on boost / libm
// (src/ContactModel.cpp:104-115, src/PoseEstimator.cpp:64-73) whose transcendental
// implementations are neither vendored nor pinned.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define ESLAM_HD __host__ __device__ __forceinline__
#else
#define ESLAM_HD static inline __attribute__((always_inline))
#endif

namespace eslam_math {

// ---------------------------------------------------------------- constants of the polynomials
// On the device a 64-bit literal costs two moves into uniform registers at every use; a __constant__ word is an operand of
// the fp64 instruction itself (c[3][..]).  The host reads the same values as literals: ESLAM_K(name).
#define ESLAM_K_LIST(X) \
    X(EXP13, 1.6059043836821613e-10) \
    X(EXP12, 2.08767569878681e-09) \
    X(EXP11, 2.505210838544172e-08) \
    X(EXP10, 2.755731922398589e-07) \
    X(EXP9, 2.7557319223985893e-06) \
    X(EXP8, 2.48015873015873e-05) \
    X(EXP7, 0.0001984126984126984) \
    X(EXP6, 0.001388888888888889) \
    X(EXP5, 0.008333333333333333) \
    X(EXP4, 0.041666666666666664) \
    X(EXP3, 0.16666666666666666) \
    X(L2E, 1.4426950408889634) \
    X(LN2_HI, 6.93147180369123816490e-01) \
    X(LN2_LO, 1.90821492927058770002e-10) \
    X(LOG21, 0.047619047619047616) \
    X(LOG19, 0.05263157894736842) \
    X(LOG17, 0.058823529411764705) \
    X(LOG15, 0.06666666666666667) \
    X(LOG13, 0.07692307692307693) \
    X(LOG11, 0.09090909090909091) \
    X(LOG9, 0.1111111111111111) \
    X(LOG7, 0.14285714285714285) \
    X(LOG5, 0.2) \
    X(LOG3, 0.3333333333333333) \
    X(SQRT2, 1.4142135623730951) \
    X(SIN19, -8.22063524662433e-18) \
    X(SIN17, 2.8114572543455206e-15) \
    X(SIN15, -7.647163731819816e-13) \
    X(SIN13, 1.6059043836821613e-10) \
    X(SIN11, -2.505210838544172e-08) \
    X(SIN9, 2.7557319223985893e-06) \
    X(SIN7, -0.0001984126984126984) \
    X(SIN5, 0.008333333333333333) \
    X(SIN3, -0.16666666666666666) \
    X(COS20, 4.110317623312165e-19) \
    X(COS18, -1.5619206968586225e-16) \
    X(COS16, 4.779477332387385e-14) \
    X(COS14, -1.1470745597729725e-11) \
    X(COS12, 2.08767569878681e-09) \
    X(COS10, -2.755731922398589e-07) \
    X(COS8, 2.48015873015873e-05) \
    X(COS6, -0.001388888888888889) \
    X(COS4, 0.041666666666666664) \
    X(TWO_PI, 6.283185307179586) \
    X(TWO_OVER_PI, 0.6366197723675814) \
    X(PIO2_1, 1.57079632673412561417e+00) \
    X(PIO2_2, 6.07710050650619224932e-11) \
    X(PIO2_3, 2.02226624879595063154e-21) \
    X(RINT_MAGIC, 6755399441055744.0)
namespace kc {
#define ESLAM_K_HOST(n, v) constexpr double h_##n = v;
ESLAM_K_LIST(ESLAM_K_HOST)
#undef ESLAM_K_HOST
#if defined(__CUDACC__)
#define ESLAM_K_DEV(n, v) static __constant__ double d_##n = v;
ESLAM_K_LIST(ESLAM_K_DEV)
#undef ESLAM_K_DEV
#endif
}
#if defined(__CUDA_ARCH__)
#define ESLAM_K(n) (::eslam_math::kc::d_##n)
#else
#define ESLAM_K(n) (::eslam_math::kc::h_##n)
#endif

// ---------------------------------------------------------------- bit moves / fma
ESLAM_HD uint64_t d2u(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; __builtin_memcpy(&u, &x, 8); return u;
#endif
}
ESLAM_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; __builtin_memcpy(&x, &u, 8); return x;
#endif
}
ESLAM_HD uint32_t f2u(float x) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u; __builtin_memcpy(&u, &x, 4); return u;
#endif
}
ESLAM_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float x; __builtin_memcpy(&x, &u, 4); return x;
#endif
}
ESLAM_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
ESLAM_HD float fmaf_(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}
ESLAM_HD double sqrt_(double x) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(x);
#else
    return __builtin_sqrt(x);
#endif
}
// round to nearest even integer, |x| < 2^51
ESLAM_HD double rint_(double x) {
    const double magic = ESLAM_K(RINT_MAGIC);  // 1.5 * 2^52
    return (x + magic) - magic;
}
ESLAM_HD bool isnan_(double x) { return x != x; }

// exact 2^k for k in [-1022, 1023]
ESLAM_HD double pow2i(int k) { return u2d((uint64_t)(k + 1023) << 52); }

// ---------------------------------------------------------------- Philox4x32-10
// Counter-based RNG (Salmon et al., SC'11).  north_star: "counter-based Philox RNG".
struct Philox4 { uint32_t v[4]; };

ESLAM_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b; hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * (uint64_t)b; lo = (uint32_t)p; hi = (uint32_t)(p >> 32);
#endif
}

ESLAM_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                               uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        philox_mulhilo(M0, c0, hi0, lo0);
        philox_mulhilo(M1, c2, hi1, lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3; return o;
}

// RNG streams (counter word 3).  Counter layout: (index, step, slot, stream); key = seed.
enum { STREAM_INIT = 0, STREAM_PROJECT = 1, STREAM_RESAMPLE = 2, STREAM_TERRAIN = 3, STREAM_HASH = 4 };

// [0,1): 53 bits, like boost::uniform_real<>(0,1) over a full-range engine
ESLAM_HD double u01(uint32_t hi, uint32_t lo) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return (double)(x >> 11) * 1.1102230246251565e-16;  // 2^-53
}
// (0,1): 52 bits + half, safe for log()
ESLAM_HD double u01_open(uint32_t hi, uint32_t lo) {
    uint64_t x = ((uint64_t)hi << 32) | lo;
    return ((double)(x >> 12) + 0.5) * 2.220446049250313e-16;  // 2^-52
}

// ---------------------------------------------------------------- exp
// exp(x) to ~1 ulp: x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor in r (tail 4e-18).
ESLAM_HD double exp_(double x) {
    if (isnan_(x)) return x;
    if (x > 709.782712893384) return u2d(0x7FF0000000000000ull);
    if (x < -745.2) return 0.0;
    const double L2E = ESLAM_K(L2E), LN2_HI = ESLAM_K(LN2_HI), LN2_LO = ESLAM_K(LN2_LO);
    double k = rint_(x * L2E);
    double r = fma_(-k, LN2_HI, x);
    r = fma_(-k, LN2_LO, r);
    double p = ESLAM_K(EXP13);               // 1/13!
    p = fma_(p, r, ESLAM_K(EXP12));            // 1/12!
    p = fma_(p, r, ESLAM_K(EXP11));            // 1/11!
    p = fma_(p, r, ESLAM_K(EXP10));            // 1/10!
    p = fma_(p, r, ESLAM_K(EXP9));            // 1/9!
    p = fma_(p, r, ESLAM_K(EXP8));            // 1/8!
    p = fma_(p, r, ESLAM_K(EXP7));            // 1/7!
    p = fma_(p, r, ESLAM_K(EXP6));            // 1/6!
    p = fma_(p, r, ESLAM_K(EXP5));            // 1/5!
    p = fma_(p, r, ESLAM_K(EXP4));            // 1/4!
    p = fma_(p, r, ESLAM_K(EXP3));            // 1/3!
    p = fma_(p, r, 0.5);
    p = fma_(p, r, 1.0);
    p = fma_(p, r, 1.0);
    int ki = (int)k;
    int k1 = ki / 2, k2 = ki - k1;                    // both in [-538, 512]
    return (p * pow2i(k1)) * pow2i(k2);
}

// ---------------------------------------------------------------- log
// log(x), x > 0 finite normal or subnormal; ~1-2 ulp.  x = 2^k m, m in [sqrt(1/2), sqrt 2),
// log m = 2 atanh(s), s = (m-1)/(m+1).
ESLAM_HD double log_(double x) {
    if (isnan_(x) || x < 0.0) return u2d(0x7FF8000000000000ull);
    if (x == 0.0) return u2d(0xFFF0000000000000ull);
    uint64_t u = d2u(x);
    int k = 0;
    if ((u >> 52) == 0) { x *= 18014398509481984.0; u = d2u(x); k = -54; }  // subnormal
    if ((u >> 52) == 0x7FF) return x;
    k += (int)(u >> 52) - 1023;
    u = (u & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;   // m in [1,2)
    double m = u2d(u);
    if (m > ESLAM_K(SQRT2)) { m *= 0.5; k += 1; }
    double f = m - 1.0;
    double s = f / (m + 1.0);
    double z = s * s;
    double p = ESLAM_K(LOG21);                  // 1/21
    p = fma_(p, z, ESLAM_K(LOG19));              // 1/19
    p = fma_(p, z, ESLAM_K(LOG17));              // 1/17
    p = fma_(p, z, ESLAM_K(LOG15));              // 1/15
    p = fma_(p, z, ESLAM_K(LOG13));              // 1/13
    p = fma_(p, z, ESLAM_K(LOG11));              // 1/11
    p = fma_(p, z, ESLAM_K(LOG9));              // 1/9
    p = fma_(p, z, ESLAM_K(LOG7));              // 1/7
    p = fma_(p, z, ESLAM_K(LOG5));              // 1/5
    p = fma_(p, z, ESLAM_K(LOG3));              // 1/3
    // log m = 2s + 2s z p
    double s2 = s + s;
    double lm = fma_(s2 * z, p, s2);
    const double LN2_HI = ESLAM_K(LN2_HI), LN2_LO = ESLAM_K(LN2_LO);
    double dk = (double)k;
    return fma_(dk, LN2_HI, fma_(dk, LN2_LO, lm));
}

// ---------------------------------------------------------------- sin / cos kernels, |a| <= pi/4
ESLAM_HD double sin_kernel(double a) {
    double z = a * a;
    double p = ESLAM_K(SIN19);                 // -1/19!
    p = fma_(p, z, ESLAM_K(SIN17));           // +-1/17!
    p = fma_(p, z, ESLAM_K(SIN15));           // +-1/15!
    p = fma_(p, z, ESLAM_K(SIN13));           // +-1/13!
    p = fma_(p, z, ESLAM_K(SIN11));           // +-1/11!
    p = fma_(p, z, ESLAM_K(SIN9));           // +-1/9!
    p = fma_(p, z, ESLAM_K(SIN7));           // +-1/7!
    p = fma_(p, z, ESLAM_K(SIN5));           // +-1/5!
    p = fma_(p, z, ESLAM_K(SIN3));           // +-1/3!
    return fma_(a * z, p, a);
}
ESLAM_HD double cos_kernel(double a) {
    double z = a * a;
    double p = ESLAM_K(COS20);                 //  1/20!
    p = fma_(p, z, ESLAM_K(COS18));          // +-1/18!
    p = fma_(p, z, ESLAM_K(COS16));          // +-1/16!
    p = fma_(p, z, ESLAM_K(COS14));          // +-1/14!
    p = fma_(p, z, ESLAM_K(COS12));          // +-1/12!
    p = fma_(p, z, ESLAM_K(COS10));          // +-1/10!
    p = fma_(p, z, ESLAM_K(COS8));          // +-1/8!
    p = fma_(p, z, ESLAM_K(COS6));          // +-1/6!
    p = fma_(p, z, ESLAM_K(COS4));          // +-1/4!
    p = fma_(p, z, -0.5);
    return fma_(z, p, 1.0);
}
ESLAM_HD void quadrant_(int q, double sk, double ck, double& s, double& c) {
    switch (q & 3) {
        case 0: s = sk;  c = ck;  break;
        case 1: s = ck;  c = -sk; break;
        case 2: s = -sk; c = -ck; break;
        default: s = -ck; c = sk; break;
    }
}
// sin/cos(2 pi u), u in [0,1): exact quadrant reduction (Box-Muller angle)
ESLAM_HD void sincos_2pi(double u, double& s, double& c) {
    double q = rint_(u * 4.0);          // 0..4
    double r = u - q * 0.25;            // exact, |r| <= 1/8
    double a = r * ESLAM_K(TWO_PI);
    quadrant_((int)q, sin_kernel(a), cos_kernel(a), s, c);
}
// sin/cos(theta), three-part Cody-Waite reduction; full accuracy for |theta| < ~1e5,
// graceful beyond (headings are un-wrapped sums of small deltas, src/PoseEstimator.cpp:206)
ESLAM_HD void sincos_(double t, double& s, double& c) {
    const double TWO_OVER_PI = ESLAM_K(TWO_OVER_PI);
    const double P1 = ESLAM_K(PIO2_1), P2 = ESLAM_K(PIO2_2), P3 = ESLAM_K(PIO2_3);
    double k = rint_(t * TWO_OVER_PI);
    double r = fma_(-k, P1, t);
    r = fma_(-k, P2, r);
    r = fma_(-k, P3, r);
    long long ki = (long long)k;
    quadrant_((int)(ki & 3), sin_kernel(r), cos_kernel(r), s, c);
}

// ---------------------------------------------------------------- normals
// Box-Muller on two open uniforms -> two independent N(0,1)
ESLAM_HD void box_muller(double u1, double u2, double& n0, double& n1) {
    double r = sqrt_(-2.0 * log_(u1));
    double s, c; sincos_2pi(u2, s, c);
    n0 = r * c; n1 = r * s;
}

// ---------------------------------------------------------------- 1/sqrt
// ~1 ulp, deterministic: fp32 magic seed (3.4e-2) + 2 fp32 Newton steps (1.8e-3, 4.7e-6) + one fp64 Halley step
// (cubic: residual 5/16 (2 * 4.7e-6)^3 = 3e-16).
// rsqrt_core is valid for x in [2^-100, 2^100); rsqrt_ adds the out-of-range fallback.
ESLAM_HD double rsqrt_core(double x) {
    // x -> float by bit moves (exponent rebias 1023 -> 127, top 23 mantissa bits: truncation) and the float result ->
    // double by bit moves (exact): both stay off the conversion unit, which is the scarcest pipe of the lean kernel
    const uint32_t xb = (uint32_t)(d2u(x) >> 29) - (896u << 23);
    const float xf = u2f(xb);
    float y = u2f(0x5f3759dfu - (xb >> 1));
    const float hx = 0.5f * xf;
    float t;
    t = y * y; t = fmaf_(-hx, t, 1.5f); y = y * t;
    t = y * y; t = fmaf_(-hx, t, 1.5f); y = y * t;
    const double yd = u2d(((uint64_t)f2u(y) << 29) + (896ull << 52));
    const double h = x * yd;
    const double e = fma_(-h, yd, 1.0);
    const double p = fma_(0.375, e, 0.5);
    return fma_(yd, e * p, yd);
}
// 2^-100 <= x < 2^100 (and positive, finite), decided on the exponent bits
ESLAM_HD bool rsqrt_in_range(double x) { return ((uint32_t)(d2u(x) >> 52) - 923u) < 200u; }
ESLAM_HD double rsqrt_(double x) {
    if (!rsqrt_in_range(x)) return 1.0 / sqrt_(x);
    return rsqrt_core(x);
}

// ---------------------------------------------------------------- inverse Mills ratio h(t) = phi(t)/Phi(t)
// The reference computes pdf(n,z)/cdf(n,z) with n = Normal(0, sigma*corr)
// (src/ContactModel.cpp:104-115) = h(z/s)/s, s = sigma*corr.  h is tabulated on
// [-H_TMAX, H_TMAX] in H_NINT intervals of width 1/H_PER_UNIT (centred on the multiples of 1/H_PER_UNIT) as degree-5
// polynomials (c0, c1 double; c2..c5 float) in the centred local coordinate (tools/gen_h_table.py; relative error <= 1e-10 everywhere a contact
// can matter, <= 1e-12 for t <= 4).  Layout: tab[q*H_NINT + j] = (c[2q], c[2q+1]) so that
// neighbouring intervals are neighbouring 16-byte words (shared-memory bank friendly).
#define ESLAM_H_TMAX 10.0
#define ESLAM_H_PER_UNIT 32
#define ESLAM_H_NINT 641
#define ESLAM_H_TABLE_ENTRIES (ESLAM_H_NINT * 2)   /* 16-byte entries */

struct dbl2 { double x, y; };

ESLAM_HD double h_slow(double t) {
    if (isnan_(t)) return t;
    if (t > 0.0) {
        // Phi(t) = 1 - O(1e-23) for t >= 10
        return exp_(-0.5 * t * t) * 0.3989422804014327;
    }
    // 1/Mills(x) = x + 1/(x + 2/(x + 3/(x + ...))), x >= 10: Laplace continued fraction
    double x = -t;
    double f = x;
    for (int k = 24; k >= 1; --k) f = x + (double)k / f;
    return f;
}

ESLAM_HD bool h_in_range(double t) { return fabs(t) < ESLAM_H_TMAX; }
// table evaluation, valid for h_in_range(t) (any other input reads a clamped interval: garbage, never out of bounds).
// Interval j is centred on t = (j - 320)/32; index and centred coordinate come from one magic-number rounding:
// m = 32 t + (1.5 2^52 + 320) has rne(32 t) + 320 in its low mantissa word, m - magic = rne(32 t) exactly,
// z = 32 t - rne(32 t) in [-0.5, 0.5] with a single rounding.  No int<->double conversions.
#define ESLAM_H_MAGIC 6755399441056064.0   /* 1.5 * 2^52 + 320 */
// interval index j (clamped) and centred coordinate z of t
ESLAM_HD void h_index(double t, uint32_t& j, double& z) {
    const double m = fma_(t, (double)ESLAM_H_PER_UNIT, ESLAM_H_MAGIC);
    j = (uint32_t)d2u(m);
    j = j > (uint32_t)(ESLAM_H_NINT - 1) ? (uint32_t)(ESLAM_H_NINT - 1) : j;   // also catches "negative" (wrapped) indices
    z = fma_(t, (double)ESLAM_H_PER_UNIT, -(m - ESLAM_H_MAGIC));
}
// c0 + c1 z + z^2 q(z), q = c2 + c3 z + c4 z^2 + c5 z^3 evaluated in fp32: with |z| <= 1/2 and intervals 1/32 wide the
// z^2 q term is below 1e-2 of the value, so its fp32 rounding (6e-8) stays below 1e-9 relative (tests/test_math.py).
// The shared-memory table is what bounds the lean kernel: 32 bytes per lookup instead of 48.
ESLAM_HD double h_poly(double z, double c0, double c1, float c2, float c3, float c4, float c5) {
    const float zf = (float)z;
    const float q = fmaf_(fmaf_(fmaf_(c5, zf, c4), zf, c3), zf, c2);
    return fma_(z * z, (double)q, fma_(c1, z, c0));
}
// layout: tab[j] = (c0, c1) as doubles, tab[H_NINT + j] = (c2, c3, c4, c5) as floats in the same 16 bytes
ESLAM_HD void h_unpack(const dbl2& hi, float& c2, float& c3, float& c4, float& c5) {
    const uint64_t a = d2u(hi.x), b = d2u(hi.y);
    c2 = u2f((uint32_t)a); c3 = u2f((uint32_t)(a >> 32)); c4 = u2f((uint32_t)b); c5 = u2f((uint32_t)(b >> 32));
}
ESLAM_HD double h_core(double t, const dbl2* tab) {
    uint32_t j; double z;
    h_index(t, j, z);
    const dbl2 c01 = tab[j];
    float c2, c3, c4, c5; h_unpack(tab[ESLAM_H_NINT + j], c2, c3, c4, c5);
    return h_poly(z, c01.x, c01.y, c2, c3, c4, c5);
}
ESLAM_HD double h_ratio(double t, const dbl2* tab) {
    if (!h_in_range(t)) return h_slow(t);
    return h_core(t, tab);
}

// contact likelihood ratio pdf/cdf for N(0, sqrt(zvar)*corr) at zdiff (src/ContactModel.cpp:104-115,170)
// inv_corr = 1.0 / contactLikelihoodCorrection
ESLAM_HD double likelihood_ratio(double zdiff, double zvar, double inv_corr, const dbl2* tab) {
    double qc = rsqrt_(zvar) * inv_corr;   // 1 / (sigma * corr)
    double t = zdiff * qc;
    return h_ratio(t, tab) * qc;
}
// the same value on the common path without branches; *rare is set when the generic function must be used instead
ESLAM_HD double likelihood_ratio_fast(double zdiff, double zvar, double inv_corr, const dbl2* tab, bool& rare) {
    const double qc = rsqrt_core(zvar) * inv_corr;
    const double t = zdiff * qc;
    rare = !(rsqrt_in_range(zvar) && h_in_range(t));
    return h_core(t, tab) * qc;
}

// ---------------------------------------------------------------- powers
// w^(1/3), w > 0 finite: w = m 2^(3q), m in [1, 8); z ~ m^(-1/3) from a quadratic seed (3e-3) and three division-free Newton steps
// (z <- z + z (1 - m z^3) / 3: error 2 e^2 each), y = m z^2, one Newton step on y with the residual y^3 - m taken exactly
// (fma).  ~1 ulp; 45 instructions against ~250 for exp(log(w) / 3): three of four contact groups found is the second most
// frequent case of pow(weight, 1/found) after four, and one lane on the generic path makes its whole warp walk it.
ESLAM_HD double cbrt_(double w) {
    uint64_t u = d2u(w);
    int bias = 0;
    if ((u >> 52) == 0) { w *= 18014398509481984.0; u = d2u(w); bias = -18; }      // subnormal: 2^54 = (2^18)^3
    const int e = (int)(u >> 52) - 1023;
    const int q = (e + 3072) / 3 - 1024, r = e - 3 * q;                            // floor(e / 3), remainder 0..2
    const double mm = u2d((u & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull);     // [1, 2)
    const double m = mm * (double)(1 << r);                                         // [1, 8), exact
    double z = fma_(fma_(0.09202394, mm, -0.47733264), mm, 1.38237822);             // mm^(-1/3) to 3e-3
    z *= r == 0 ? 1.0 : (r == 1 ? 0.7937005259840998 : 0.6299605249474366);         // 2^(-r/3)
    const double third = 0.3333333333333333;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int it = 0; it < 3; ++it) {
        const double t = (z * z) * z;
        const double h = fma_(-m, t, 1.0);
        z = fma_(z * third, h, z);
    }
    double y = (m * z) * z;
    const double y2 = y * y, e2 = fma_(y, y, -y2);                                  // y^2 = y2 + e2 exactly
    const double res = fma_(e2, y, fma_(y2, y, -m));                                // y^3 - m
    y = fma_(-(res * third), z * z, y);                                             // 1 / (3 y^2) = z^2 / 3
    return y * pow2i(q + bias);
}

// w^(1/n) as in pow(weight, 1.0/found_points) (src/PoseEstimator.cpp:309), w >= 0
ESLAM_HD double root_n(double w, unsigned long long n) {
    if (n == 1) return w;
    if (n == 2) return sqrt_(w);
    if (n == 4) return sqrt_(sqrt_(w));
    if (w == 0.0) return 0.0;
    if (n == 3 && w < 1.7e308) return cbrt_(w);                                     // finite (NaN / inf: the generic path)
    if (n == 0) return (w == 1.0) ? 1.0 : (w < 1.0 ? 0.0 : u2d(0x7FF0000000000000ull));  // pow(w, inf)
    return exp_(log_(w) / (double)n);
}
// x^e for the size_t exponent 4 - cpoints.size() (src/PoseEstimator.cpp:336), x >= 0.
// e wraps to ~1.8e19 when more than 4 contacts were found (quirk Q2).
ESLAM_HD double pow_u64(double x, unsigned long long e) {
    if (e <= 16) {
        double r = 1.0;
        for (unsigned long long i = 0; i < e; ++i) r *= x;
        return r;
    }
    if (isnan_(x)) return x;
    if (x == 1.0) return 1.0;
    return (x < 1.0) ? 0.0 : u2d(0x7FF0000000000000ull);
}

// ---------------------------------------------------------------- grid cell index
// floor((x - offset) / scale) as envire's GridBase::toGrid does it (un-vendored).
// Returns false when outside [0, n).  The device fast path multiplies by a reciprocal and only
// falls back to this exact division when the quotient is within 2^-18 of an integer.
ESLAM_HD bool cell_index_exact(double x, double offset, double scale, int n, int& idx) {
    double q = floor((x - offset) / scale);
    if (!(q >= 0.0 && q < (double)n)) return false;
    idx = (int)q; return true;
}

// ---------------------------------------------------------------- SurfaceHash (src/SurfaceHash.hpp)
// SurfaceParam::fromPoints (:60-110): least-squares plane z = sx x + sy y + c through n >= 3 points; the normal equations
// [[xx xy x] [xy yy y] [x y n]] r = (xz yz z) solved by LDL^T (the reference calls Eigen's pivoting LDLT: the same solution up
// to rounding -- Eigen is not vendored).  Returns false for a singular system (collinear points).
ESLAM_HD bool surface_slopes(const double* px, const double* py, const double* pz, int n, double& slope_x, double& slope_y) {
    double x = 0, y = 0, z = 0, xx = 0, yy = 0, xy = 0, xz = 0, yz = 0;
    for (int i = 0; i < n; ++i) {
        x += px[i]; y += py[i]; z += pz[i];
        xx += px[i] * px[i]; yy += py[i] * py[i]; xy += px[i] * py[i]; xz += px[i] * pz[i]; yz += py[i] * pz[i];
    }
    const double d0 = xx;
    if (!(d0 > 0.0)) return false;
    const double l10 = xy / d0, l20 = x / d0;
    const double d1 = yy - l10 * xy;
    if (!(d1 > 0.0)) return false;
    const double l21 = (y - l20 * xy) / d1;
    const double d2 = ((double)n - l20 * x) - l21 * l21 * d1;
    if (!(d2 > 0.0)) return false;
    // forward: L w = b;  diagonal;  backward: L^T r = w / d
    const double w0 = xz, w1 = yz - l10 * w0, w2 = (z - l20 * w0) - l21 * w1;
    const double r2 = w2 / d2;
    const double r1 = w1 / d1 - l21 * r2;
    const double r0 = (w0 / d0 - l10 * r1) - l20 * r2;
    slope_x = r0; slope_y = r1;
    return true;
}
// Buckets::bucketIndex (:26-30) for the slope hash: count buckets over [-1, 1]; NaN -> bucket 0
ESLAM_HD int slope_bucket(double value, int count) {
    const double v = (value - (-1.0)) / (1.0 - (-1.0)) * (double)count;
    if (!(v > 0.0)) return 0;
    if (v >= (double)count) return count - 1;
    return (int)v;
}
// the 4-point footprint of SurfaceHash::create (:162-168): base 0.5 m
ESLAM_HD void hash_footprint(int i, double& x, double& y) {
    const double b = 0.5;
    x = (i & 1) ? -b / 2.0 : b / 2.0;
    y = (i & 2) ? -b : 0.0;
}

}  // namespace eslam_math
