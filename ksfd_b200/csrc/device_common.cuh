// Device-side data structures and pointwise physics shared by all kernels.
//
// Layout of every field vector: fp64, dof fastest, then x, y, z (reference
// KSFD/ksfdgrid.py:10-28).  A rank owns `nloc` planes of the LAST spatial
// axis; the two ghost planes on either side are reached through VecRef.lo /
// VecRef.hi (on one rank they alias the vector itself: periodic wrap).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/ksfd_b200.h"

#define KSFD_SW 2                       // stencil width (order 3)
#define KSFD_RING 6                     // ring slots of the marching kernels

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
// Pointwise free energy G(rho, U) = V + s2*log(rho)
// (KSFD/ksfdsym.py:983-990; KSFD/ksfdligand.py:527-547; ksfdsoln.py:147-161)
// ---------------------------------------------------------------------------
template <int NLIG>
__device__ __forceinline__ double G_point(const DevPhys &P, double rho,
                                          const double *U)
{
    double G = P.s2 * log(rho);
    for (int g = 0; g < P.ngroups; ++g) {
        double sU = 0.0;
#pragma unroll
        for (int l = 0; l < NLIG; ++l)
            if (P.lig_group[l] == g) sU = fma(P.weight[l], U[l], sU);
        G = fma(-P.beta[g], log(P.alpha[g] + sU), G);
    }
    double th = tanh((rho - P.rhomax) * P.inv_cushion);
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
