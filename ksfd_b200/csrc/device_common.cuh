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
#include <math_constants.h>
#include "../../include/ksfd_b200.h"
#include "fastmath.cuh"

// Flags of the pipelined solver (skip / cycle state) are plain device words.
// (Programmatic dependent launch of the Krylov-loop kernels — griddepcontrol.wait +
// launch_dependents at kernel entry, flags read from L2 — was built and measured in round 2:
// 2.229 -> 2.192 ms per 1024^2 step on one B200, 2.952 -> 3.022 ms on two
// (profiles/r02_pdl_ab.txt); not worth a second code path, removed.)
#define KSFD_FLAG(p) (*(p))
#define KSFD_PDL_ENTER() do { } while (0)

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
    // dense group x ligand weights (0 where the ligand is not in the group):
    // sum_l Wgl[g][l]*U_l needs no membership test in the inner loops
    double Wgl[KSFD_MAX_GROUPS][KSFD_MAX_LIGANDS];
    double w2c;                 // sum over axes of the centre weight of d2/dx2
    // symmetric form of the 4th-order stencils (marching kernels):
    //   d/dx   = c1 * (8 (f[+1] - f[-1]) - (f[+2] - f[-2]))
    //   d2/dx2 = c2 * (16 (f[+1] + f[-1]) - (f[+2] + f[-2]) - 30 f[0])
    // c1 = w1[+1]/8, c2 = w2[-1]/16 (exact rescalings of reference weights); the
    // other reference weights agree with this pattern to a few ulp (sym_ok,
    // checked on the host; otherwise the direct kernels run)
    double c1[3], c2[3], c1sq[3];
    int sym_ok, pad_;
    // cap potential through the logistic: tanh(x)+1 = 2/(1+exp(-2x)),
    // -2x = ycap0 + ycap1*rho ;  cap = capscale2 / (1 + exp(-2x))
    double ycap0, ycap1, capscale2;
    FastK mk;                   // log/exp constants (fastmath.cuh)
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
    // peer-to-peer halos are double-buffered on the parity of a DEVICE-side
    // exchange counter: the ghost planes live at lo/hi + (*par & 1) * pstride
    const unsigned long long *par;
    long long pstride;
    // Several ranks over NVLink peer memory: the neighbours store the ghost planes and then
    // publish the exchange number in flag_lo / flag_hi.  The exchange kernels only PUSH;
    // the marching kernels wait here, in the CTAs that actually read ghost planes and
    // right before they do (march_kernels.cuh: plane_of, tma_march.cuh: tma_halo_wait),
    // so the NVLink round trip hides behind the owned planes.  nullptr: nothing to wait for.
    const volatile unsigned long long *flag_lo, *flag_hi;
    volatile int *err;                      // host-visible: a wait timed out
    volatile unsigned long long *dead;      // device-side sticky copy of it
};

// Flag words of the peer-memory exchanges.  The consumer ACQUIRES the flag (ld.acquire.sys:
// what it reads afterwards is ordered behind the observation, no full fence — a
// system-scope MEMBAR in an SM that is streaming stores waits for all of them, measured
// 10-20 us inside the bandwidth-bound kernels), the producer RELEASES it (st.release.sys).
#ifndef KSFD_FLAG_ACQREL
#define KSFD_FLAG_ACQREL 1
#endif
__device__ __forceinline__ unsigned long long flag_acquire(const volatile unsigned long long *f)
{
#if KSFD_FLAG_ACQREL
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
    return v;
#else
    return *f;
#endif
}
// after a successful wait
__device__ __forceinline__ void flag_acquired()
{
#if !KSFD_FLAG_ACQREL
    __threadfence_system();
#endif
}
__device__ __forceinline__ void flag_release(volatile unsigned long long *f, unsigned long long v)
{
#if KSFD_FLAG_ACQREL
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(v) : "memory");
#else
    __threadfence_system();
    *f = v;
#endif
}

// bounded spin until the neighbour's flag reaches exchange number q (cf. p2p_spin)
static __device__ __noinline__ void halo_flag_wait(const volatile unsigned long long *f,
                                                   unsigned long long q, volatile int *err,
                                                   volatile unsigned long long *dead)
{
    const long long t0 = clock64();
    unsigned spins = 0;
    while (flag_acquire(f) < q) {
        __nanosleep(20);
        if ((++spins & 0xfff) == 0 && (clock64() - t0 > 240000000000ll || (dead && *dead))) {
            if (dead) *dead = 1ull;
            if (err) *err = 1;
            break;
        }
    }
    flag_acquired();
}

__device__ __forceinline__ long long ghost_shift(const VecRef &v)
{
    return v.par ? (long long)(*v.par & 1ull) * v.pstride : 0;
}

__device__ __forceinline__ const double *plane_ptr(const VecRef &v, int k,
                                                   int nloc, long long stride)
{
    if (k < 0) return v.lo + ghost_shift(v) + (long long)(k + KSFD_SW) * stride;
    if (k >= nloc) return v.hi + ghost_shift(v) + (long long)(k - nloc) * stride;
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
    return x > lo ? x : lo;     // NaN > lo is false -> lo
}

// ---------------------------------------------------------------------------
// Pointwise free energy G(rho, U) = V + s2*log(rho), libm version for the
// direct kernels and the Jacobian set-up (the marching kernels use the
// table-driven G_fast of march_kernels.cuh)
// (KSFD/ksfdsym.py:983-990; KSFD/ksfdligand.py:527-547; ksfdsoln.py:147-161)
// ---------------------------------------------------------------------------
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
