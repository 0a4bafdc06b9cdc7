// Device-side data structures and pointwise physics shared by all kernels.
//
// INTERNAL device layout of every field vector ("plane-SoA"): for each plane
// k of the LAST spatial axis, the nf fields of that plane are stored one after
// the other, each as plane_pts contiguous fp64 values (x fastest, then y):
//     element(k, c, pp) = ((k * nf + c) * plane_pts + pp)
// so that a warp reading one field of consecutive x points issues one fully
// coalesced 256-byte request, and the two ghost planes of a slab are still
// one contiguous block.  The reference layout (dof fastest, then x, y, z;
// KSFD/ksfdgrid.py:10-28) exists only at the boundary: ksfd_to_internal /
// ksfd_from_internal convert (in 1-D the two layouts coincide).
// A rank owns `nloc` planes; the two ghost planes on either side are reached
// through VecRef.lo / VecRef.hi (on one rank they alias the vector itself:
// periodic wrap).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/ksfd_b200.h"

#define KSFD_SW 2                       // stencil width (order 3)
#define KSFD_RING 4                     // smem ring slots of the marching kernels

struct DevPhys {
    int ngroups, nlig, cap_type, dim;
    double s2, rhomax, inv_cushion, capscale, rhomin, Umin, inv_rhomax;
    double alpha[KSFD_MAX_GROUPS], beta[KSFD_MAX_GROUPS];
    int lig_group[KSFD_MAX_LIGANDS + 1];
    double weight[KSFD_MAX_LIGANDS], s[KSFD_MAX_LIGANDS];
    double gamma[KSFD_MAX_LIGANDS], D[KSFD_MAX_LIGANDS];
    double w1[3][5], w2[3][5];
};

struct Geom {
    int dim, dof;
    int n0, n1;                 // extents of the non-marching axes (x, y)
    int nloc;                   // owned planes along the last axis
    long long plane_pts;        // points in one plane of the last axis
    long long npts;             // owned points
};

// A vector plus its ghost planes along the last axis.
struct VecRef {
    const double *base;         // plane 0 .. nloc-1
    const double *lo;           // planes -2, -1
    const double *hi;           // planes nloc, nloc+1
};

__device__ __forceinline__ const double *plane_ptr(const VecRef &v, int k,
                                                   int nloc, long long stride)
{
    if (k < 0) return v.lo + (long long)(k + KSFD_SW) * stride;
    if (k >= nloc) return v.hi + (long long)(k - nloc) * stride;
    return v.base + (long long)k * stride;
}

__device__ __forceinline__ int wrapi(int i, int n)
{
    i %= n;
    return i < 0 ? i + n : i;
}

// clamp as the reference's groom: max(x, lo); NaN -> lo (fmax returns the
// non-NaN operand)  (KSFD/ksfdsym.py:888-900)
__device__ __forceinline__ double clampv(double x, double lo)
{
    return fmax(x, lo);
}

// ---------------------------------------------------------------------------
// fp64 log / tanh with their polynomial coefficients in the constant bank
// (DFMA takes c[bank][off] operands directly; CUDA's libm materialises every
// coefficient with two UMOVs, ~110 extra issue slots per G evaluation).
// Accuracy: < 1 ulp (log, fdlibm __ieee754_log scheme) and < 2 ulp (tanh via
// expm1-free exp), same class as the libm routines they replace; arguments
// outside the fast range fall back to libm.
// ---------------------------------------------------------------------------
__constant__ double c_log[10] = {
    6.666666666666735130e-01, 3.999999999940941908e-01, 2.857142874366239149e-01,
    2.222219843214978396e-01, 1.818357216161805012e-01, 1.531383769920937332e-01,
    1.479819860511658591e-01,
    6.93147180369123816490e-01,      // ln2_hi
    1.90821492927058770002e-10,      // ln2_lo
    0.0};

// n/d for finite normal operands of moderate magnitude (no overflow/underflow
// handling): MUFU.RCP64H seed + one Newton step + one residual correction.
__device__ __forceinline__ double fast_div(double n, double d)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    // seed ~2^-20 -> one Newton step ~2^-40 -> quotient with residual
    // correction: error ~2^-80 before the final rounding
    const double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    const double q = n * r;
    return fma(fma(-d, q, n), r, q);
}

__device__ __forceinline__ double fast_log(double x)
{
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    // fast path: positive, normal, finite; everything else goes to libm
    if ((unsigned)(hx - 0x00100000) >= 0x7fe00000u) return log(x);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    const double m = __hiloint2double(hx | (i ^ 0x3ff00000), lx);   // [sqrt2/2, sqrt2)
    k += (i >> 20);
    const double f = m - 1.0;
    const double s = fast_div(f, 2.0 + f);
    const double dk = (double)k;
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * fma(w, fma(w, c_log[5], c_log[3]), c_log[1]);
    const double t2 = z * fma(w, fma(w, fma(w, c_log[6], c_log[4]), c_log[2]), c_log[0]);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    return fma(dk, c_log[7], -((hfsq - fma(s, hfsq + R, dk * c_log[8])) - f));
}

__constant__ double c_exp[14] = {
    1.4426950408889634e+00,          // 1/ln2
    6.93147180369123816490e-01,      // ln2_hi
    1.90821492927058770002e-10,      // ln2_lo
    // exp(r) - 1 - r = r^2 * (1/2! + r/3! + ...), |r| <= ln2/2
    5.0000000000000000e-01, 1.6666666666666666e-01, 4.1666666666666664e-02,
    8.3333333333333332e-03, 1.3888888888888889e-03, 1.9841269841269841e-04,
    2.4801587301587302e-05, 2.7557319223985893e-06, 2.7557319223985888e-07,
    2.5052108385441720e-08, 2.0876756987868100e-09};

// tanh(x) = expm1(2|x|) / (expm1(2|x|) + 2), sign restored; expm1 is formed
// without cancellation as 2^k*p + (2^k - 1) with p = exp(r) - 1.
__device__ __forceinline__ double fast_tanh(double x)
{
    const double a = fabs(x);
    if (!(a < 20.0)) return (a != a) ? x : copysign(1.0, x);
    const double y = a + a;
    const double kd = rint(y * c_exp[0]);
    const int k = (int)kd;
    const double r = fma(-kd, c_exp[2], fma(-kd, c_exp[1], y));
    // P(r) = sum_{i=0}^{10} c[3+i] r^i by Estrin's scheme (short dependency chains)
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double p01 = fma(c_exp[4], r, c_exp[3]);
    const double p23 = fma(c_exp[6], r, c_exp[5]);
    const double p45 = fma(c_exp[8], r, c_exp[7]);
    const double p67 = fma(c_exp[10], r, c_exp[9]);
    const double p89 = fma(c_exp[12], r, c_exp[11]);
    const double p03 = fma(p23, r2, p01);
    const double p47 = fma(p67, r2, p45);
    const double p8a = fma(c_exp[13], r2, p89);
    double p = fma(p8a, r8, fma(p47, r4, p03));
    p = fma(p * r, r, r);                                 // exp(r) - 1
    const double s2k = __hiloint2double((1023 + k) << 20, 0);   // 2^k, k in [0, 58]
    const double em1 = fma(s2k, p, s2k - 1.0);
    return copysign(fast_div(em1, em1 + 2.0), x);
}

// ---------------------------------------------------------------------------
// Pointwise free energy G(rho, U) = V + s2*log(rho)
// (KSFD/ksfdsym.py:983-990; KSFD/ksfdligand.py:527-547; ksfdsoln.py:147-161)
// ---------------------------------------------------------------------------
template <int NLIG>
__device__ __forceinline__ double G_point(const DevPhys &P, double rho,
                                          const double *U)
{
    // straight-line code (groups unrolled: ngroups <= NLIG) so that the
    // independent log/tanh evaluations interleave and hide fp64 latency
    double G = P.s2 * fast_log(rho);
    const double th = fast_tanh((rho - P.rhomax) * P.inv_cushion);
#pragma unroll
    for (int g = 0; g < NLIG; ++g) {
        if (g < P.ngroups) {
            double sU = 0.0;
#pragma unroll
            for (int l = 0; l < NLIG; ++l)
                if (P.lig_group[l] == g) sU = fma(P.weight[l], U[l], sU);
            G = fma(-P.beta[g], fast_log(P.alpha[g] + sU), G);
        }
    }
    double cap = P.capscale * (th + 1.0);
    if (P.cap_type == 1) cap *= rho * P.inv_rhomax;
    return G + cap;
}

// runtime-nlig variant for the generic (naive) kernels
__device__ __forceinline__ double G_point_rt(const DevPhys &P, double rho,
                                             const double *U)
{
    double G = P.s2 * log(rho);
    for (int g = 0; g < P.ngroups; ++g) {
        double sU = 0.0;
        for (int l = 0; l < P.nlig; ++l)
            if (P.lig_group[l] == g) sU = fma(P.weight[l], U[l], sU);
        G = fma(-P.beta[g], log(P.alpha[g] + sU), G);
    }
    double th = tanh((rho - P.rhomax) * P.inv_cushion);
    double cap = P.capscale * (th + 1.0);
    if (P.cap_type == 1) cap *= rho * P.inv_rhomax;
    return G + cap;
}

// G and its partials dG/drho, dG/dU_l (chain rule through log/tanh; what the
// reference obtains symbolically, KSFD/ksfdsym.py:1021-1033,1094-1100)
__device__ __forceinline__ void G_and_partials_rt(const DevPhys &P, double rho,
                                                  const double *U, double &G,
                                                  double &g_rho, double *g_U)
{
    G = P.s2 * log(rho);
    for (int g = 0; g < P.ngroups; ++g) {
        double sU = 0.0;
        for (int l = 0; l < P.nlig; ++l)
            if (P.lig_group[l] == g) sU = fma(P.weight[l], U[l], sU);
        double a = P.alpha[g] + sU;
        G = fma(-P.beta[g], log(a), G);
        double ia = 1.0 / a;
        for (int l = 0; l < P.nlig; ++l)
            if (P.lig_group[l] == g) g_U[l] = -P.beta[g] * P.weight[l] * ia;
    }
    double th = tanh((rho - P.rhomax) * P.inv_cushion);
    double sech2 = 1.0 - th * th;
    double cap = P.capscale * (th + 1.0);
    double dcap;
    if (P.cap_type == 1) {
        dcap = P.capscale * (sech2 * P.inv_cushion * (rho * P.inv_rhomax) +
                             (th + 1.0) * P.inv_rhomax);
        cap *= rho * P.inv_rhomax;
    } else {
        dcap = P.capscale * sech2 * P.inv_cushion;
    }
    G += cap;
    g_rho = P.s2 / rho + dcap;
}

// block reductions -----------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}
// non-negative doubles order like their bit patterns
__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v)
{
    atomicMax(reinterpret_cast<unsigned long long *>(addr),
              static_cast<unsigned long long>(__double_as_longlong(v)));
}
