// Table-driven fp64 log and logistic (1/(1+exp)) for the stencil kernels.
//
// Why not libm / the polynomial-only routines: on sm_100 fp64 instructions take
// constants from uniform registers only (63 of them), so every polynomial
// coefficient competes with the physics and stencil constants; and the
// kernels are issue-bound on the transcendentals.  A 128-entry (1/c, log c)
// table brings log down to 10 fp64 operations with 5 constants, a 64-entry
// 2^(j/64) table brings exp to 10 with 6.  The tables (2.5 KB) are staged in
// shared memory once per CTA.
//
// Accuracy (tests/test_fastmath.py compiles this header for the host and
// compares with long double): log: |err| <= 1.25 ulp(result) + 2^-58 absolute
// (the absolute term only matters for |log x| < 1e-2, where the table value and
// the series cancel; log(1) = 1.7e-18); exp: <= 1.5 ulp; logistic: <= 4 ulp.
// Same class as the
// libm routines they replace (the reference links glibc; agreement with it is
// to rounding either way).
#pragma once
#include "fastmath_tables.h"

#ifdef __CUDACC__
#define KSFD_HD __host__ __device__ __forceinline__
#else
#define KSFD_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define KSFD_HI(x) __double2hiint(x)
#define KSFD_LO(x) __double2loint(x)
#define KSFD_HILO(h, l) __hiloint2double(h, l)
#define KSFD_FMA(a, b, c) fma(a, b, c)
#else
#include <cmath>
#include <cstring>
static inline int ksfd_hi_(double x)
{
    unsigned long long b;
    std::memcpy(&b, &x, 8);
    return (int)(b >> 32);
}
static inline int ksfd_lo_(double x)
{
    unsigned long long b;
    std::memcpy(&b, &x, 8);
    return (int)(b & 0xffffffffu);
}
static inline double ksfd_hilo_(int h, int l)
{
    unsigned long long b = ((unsigned long long)(unsigned)h << 32) | (unsigned)l;
    double x;
    std::memcpy(&x, &b, 8);
    return x;
}
#define KSFD_HI(x) ksfd_hi_(x)
#define KSFD_LO(x) ksfd_lo_(x)
#define KSFD_HILO(h, l) ksfd_hilo_(h, l)
#define KSFD_FMA(a, b, c) std::fma(a, b, c)
#endif

// table access: host/test = plain arrays; device = the CTA's shared-memory copy
struct FastTabs {
    const double *logt;     // [128][2] = (invc, logc)
    const double *expt;     // [64]
    KSFD_HD void log_pair(int i, double &invc, double &logc) const
    {
        invc = logt[2 * i];
        logc = logt[2 * i + 1];
    }
    KSFD_HD double exp2j(int j) const { return expt[j]; }
};

// polynomial / range-reduction constants.  They travel in the kernel parameter
// block: a literal in the code costs two move instructions per use, a
// parameter is one uniform load (or stays in a uniform register).
struct FastK {
    double l_a1, l_a3, l_a4;            // 1/3, 1/5, -1/6   (log1p series; -1/2, -1/4 are immediates)
    double ln2_hi, ln2_lo;
    double e_inv, e_hi, e_lo;           // 64/ln2, ln2/64 split
    double e_c3, e_c4, e_c5;            // 1/6, 1/24, 1/120
};

inline FastK fastk_default()
{
    FastK k;
    k.l_a1 = 0x1.5555555555555p-2;
    k.l_a3 = 0x1.999999999999ap-3;
    k.l_a4 = -0x1.5555555555555p-3;
    k.ln2_hi = KSFD_LN2_HI;
    k.ln2_lo = KSFD_LN2_LO;
    k.e_inv = KSFD_64_LN2;
    k.e_hi = KSFD_LN2_64_HI;
    k.e_lo = KSFD_LN2_64_LO;
    k.e_c3 = 0x1.5555555555555p-3;
    k.e_c4 = 0x1.5555555555555p-5;
    k.e_c5 = 0x1.1111111111111p-7;
    return k;
}

// reciprocal of a positive normal double: hardware seed (20+ bits) and two
// Newton steps (error ~2^-80 before the last rounding)
KSFD_HD double ksfd_rcp(double d)
{
    double r;
#if defined(__CUDA_ARCH__)
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
#else
    r = KSFD_HILO(KSFD_HI(1.0 / d) & ~0x7, 0);          // ~20-bit seed like MUFU.RCP64H
#endif
    double e = KSFD_FMA(-d, r, 1.0);
    r = KSFD_FMA(r, e, r);
    e = KSFD_FMA(-d, r, 1.0);
    return KSFD_FMA(r, e, r);
}

// true iff x is positive, normal and finite (the domain of ksfd_log_core)
KSFD_HD bool ksfd_log_domain(double x)
{
    return (unsigned)(KSFD_HI(x) - 0x00100000) < 0x7fe00000u;
}

// log x for positive normal finite x
template <class TA>
KSFD_HD double ksfd_log_core(double x, const FastK &K, const TA &T, int kadj = 0)
{
    const int hx = KSFD_HI(x);
    const int tmp = hx - 0x3fe60000;
    const int i = (tmp >> 13) & 127;
    const int k = (tmp >> 20) + kadj;                   // arithmetic shift
    const double z = KSFD_HILO(hx - (tmp & (int)0xfff00000), KSFD_LO(x));   // [0.6875, 1.375)
    double invc, logc;
    T.log_pair(i, invc, logc);
    const double r = KSFD_FMA(z, invc, -1.0);           // exact up to one rounding
    const double kd = (double)k;
    const double w = KSFD_FMA(kd, K.ln2_hi, logc);
    const double r2 = r * r;
    // log1p(r) = r + r^2 (-1/2 + r/3 - r^2/4 + r^3/5 - r^4/6),  |r| < 2^-8
    const double p01 = KSFD_FMA(r, K.l_a1, -0.5);
    const double p23 = KSFD_FMA(r, K.l_a3, -0.25);
    const double p = KSFD_FMA(r2, KSFD_FMA(r2, K.l_a4, p23), p01);
    const double t = KSFD_FMA(r2, p, KSFD_FMA(kd, K.ln2_lo, r));
    return w + t;
}

// log x for any x: NaN/+inf -> x, negative -> NaN, 0 -> -inf, subnormal ->
// rescaled by 2^54
template <class TA>
KSFD_HD double ksfd_log(double x, const FastK &K, const TA &T)
{
    if (ksfd_log_domain(x)) return ksfd_log_core(x, K, T);
    if (!(x < 1.79769313486231570815e+308)) return x + x;      // NaN, +inf
    if (x < 0.0) return KSFD_HILO(0x7ff80000, 0);
    if (x == 0.0) return KSFD_HILO((int)0xfff00000, 0);
    return ksfd_log_core(x * 18014398509481984.0, K, T, -54);
}

// exp(y) for |y| <= 707
template <class TA>
KSFD_HD double ksfd_exp_core(double y, const FastK &K, const TA &T)
{
    const double SHIFT = 6755399441055744.0;            // 1.5 * 2^52
    const double zs = KSFD_FMA(y, K.e_inv, SHIFT);
    const int ki = KSFD_LO(zs);                          // round(y*64/ln2)
    const double kd = zs - SHIFT;
    double r = KSFD_FMA(kd, -K.e_hi, y);
    r = KSFD_FMA(kd, -K.e_lo, r);                        // |r| <= ln2/128
    const double tj = T.exp2j(ki & 63);
    // exp(r) - 1 = r + r^2 (1/2 + r/6 + r^2/24 + r^3/120)
    const double r2 = r * r;
    const double p01 = KSFD_FMA(r, K.e_c3, 0.5);
    const double p23 = KSFD_FMA(r, K.e_c5, K.e_c4);
    const double p = KSFD_FMA(r2, KSFD_FMA(r2, p23, p01), r);
    const double e0 = KSFD_FMA(tj, p, tj);               // [0.99, 2.02)
    return KSFD_HILO(KSFD_HI(e0) + ((ki >> 6) << 20), KSFD_LO(e0));
}

template <class TA>
KSFD_HD double ksfd_exp(double y, const FastK &K, const TA &T)
{
    y = y < 707.0 ? y : 707.0;                           // NaN -> 707
    y = y > -707.0 ? y : -707.0;
    return ksfd_exp_core(y, K, T);
}

// 1/(1+exp(y)):  tanh(x) + 1 = 2/(1+exp(-2x)) without the cancellation of
// computing tanh first
template <class TA>
KSFD_HD double ksfd_logistic(double y, const FastK &K, const TA &T)
{
    return ksfd_rcp(1.0 + ksfd_exp(y, K, T));
}
