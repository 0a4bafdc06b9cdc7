// Marching stencil kernels (the performance path, 2-D and 3-D).
//
// A CTA owns a tile of TXL x TYL "lanes" (one thread each) in the non-marching
// axes (x in 2-D, x*y in 3-D; two halo lanes on every side, periodic wrap) and
// marches along the LAST axis over RZ output planes.  Per plane:
//   stage: the plane's points (plane-SoA => fully coalesced loads, software-
//          prefetched one plane ahead) are clamped and turned into the
//          pointwise fields the stencil needs — for the residual: rho,
//          G(rho,U) [3 log + 1 tanh, evaluated ONCE per point], U_l.  Each
//          thread pushes its lane's fields into a 5-deep REGISTER queue (the
//          marching-axis stencil never touches memory) and into one slot of a
//          4-slot shared-memory ring (for the cross-axis neighbours);
//   emit : after ONE __syncthreads, the 4th-order star stencil of the plane two
//          behind is formed from the queue (marching axis) and from
//          neighbouring lanes of the ring slot (cross axes: immediate-offset
//          LDS, strides are compile-time) and the outputs are stored coalesced.
// The field vector is read once from HBM (halo re-reads hit L2), G is evaluated
// (TXL*TYL/out) * (RZ+4)/RZ times per output point instead of 1+4*dim times,
// and no ghosted copy or G array is ever materialised.
//
// The skeleton is shared by three operators (Op policies): residual, J.v
// (optionally fused with the block-Jacobi solve), velocity(-max).
#pragma once
#include "device_common.cuh"
#include "naive_kernels.cuh"   // pc_point

// stencil access for one lane at emit time
template <int DIM, int NF, int TXL, int NT>
struct RegAcc {
    const double (*q)[5];       // q[f][0..4] = planes ko-2 .. ko+2 of this lane
    const double *rc;           // &ring[centre slot][0][tid]
    __device__ __forceinline__ double c(int f) const { return q[f][2]; }
    // weighted 5-point sum of field f along axis ax with weights w[5]
    __device__ __forceinline__ double wsum(int f, int ax, const double *w) const
    {
        double r;
        if (ax == DIM - 1) {
            r = w[0] * q[f][0];
            r = fma(w[1], q[f][1], r);
            r = fma(w[2], q[f][2], r);
            r = fma(w[3], q[f][3], r);
            r = fma(w[4], q[f][4], r);
        } else {
            const int st = (ax == 0) ? 1 : TXL;
            const double *b = rc + f * NT;
            r = w[0] * b[-2 * st];
            r = fma(w[1], b[-st], r);
            r = fma(w[2], q[f][2], r);
            r = fma(w[3], b[st], r);
            r = fma(w[4], b[2 * st], r);
        }
        return r;
    }
};

// ---------------------------------------------------------------------------
// Operator policies
// ---------------------------------------------------------------------------
// Residual: F = udot - (f(u)+src)  |  f(u)+src
// (KSFD/ksfdsym.py:902-940 + KSFD/ksfdts.py:591-592 fused)
template <int DIM, int NLIG>
struct ResidualOp {
    static constexpr int NF = NLIG + 2;        // rho, G, U_l
    static constexpr int NPRE = NLIG + 1;
    static constexpr int NAUX = NLIG + 1;      // udot of the output plane
    VecRef u;
    const double *udot, *src;
    double *out;
    struct State {};

    __device__ __forceinline__ void load(const Geom &g, int k, long long poff,
                                         double *pre) const
    {
        const double *p = plane_ptr(u, k, g.nloc, g.plane_pts * (NLIG + 1)) + poff;
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) pre[c] = __ldg(p + c * g.plane_pts);
    }
    __device__ __forceinline__ void load_aux(const Geom &g, int ko, long long poff,
                                             double *aux) const
    {
        if (udot) {
            const double *p = udot + (long long)ko * (NLIG + 1) * g.plane_pts + poff;
#pragma unroll
            for (int c = 0; c < NLIG + 1; ++c) aux[c] = __ldg(p + c * g.plane_pts);
        }
    }
    __device__ __forceinline__ void stage(const DevPhys &P, const double *pre,
                                          double *f) const
    {
        double rho = clampv(pre[0], P.rhomin);
        double U[NLIG];
#pragma unroll
        for (int l = 0; l < NLIG; ++l) U[l] = clampv(pre[1 + l], P.Umin);
        f[0] = rho;
        f[1] = G_point<NLIG>(P, rho, U);
#pragma unroll
        for (int l = 0; l < NLIG; ++l) f[2 + l] = U[l];
    }
    template <class Acc>
    __device__ __forceinline__ void emit(const DevPhys &P, const Geom &g, int ko,
                                         long long poff, const Acc &a,
                                         const double *aux, State &) const
    {
        const double rho0 = a.c(0);
        double acc = 0.0, lap = 0.0;
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) {
            acc = fma(a.wsum(0, ax, P.w1[ax]), a.wsum(1, ax, P.w1[ax]), acc);
            lap += a.wsum(1, ax, P.w2[ax]);
        }
        double f[NLIG + 1];
        f[0] = fma(rho0, lap, acc);
#pragma unroll
        for (int l = 0; l < NLIG; ++l) {
            double lapU = 0.0;
#pragma unroll
            for (int ax = 0; ax < DIM; ++ax) lapU += a.wsum(2 + l, ax, P.w2[ax]);
            f[1 + l] =
                fma(P.D[l], lapU, fma(P.s[l], rho0, -P.gamma[l] * a.c(2 + l)));
        }
        const long long e = (long long)ko * (NLIG + 1) * g.plane_pts + poff;
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) {
            double v = f[c];
            if (src) v += __ldg(src + e + c * g.plane_pts);
            out[e + c * g.plane_pts] = udot ? aux[c] - v : v;
        }
    }
    __device__ __forceinline__ void finish(State &) const {}
};

// J.v: out = (shift*I - J(u_lin)) * z,  z = v or M^{-1} v
// (replaces the assembled matrix of KSFD/ksfdsym.py:814-886)
template <int DIM, int NLIG, bool PRECOND>
struct JvpOp {
    static constexpr int NF = NLIG + 4;   // z_rho, dG, z_U.., rho, G
    static constexpr int NPRE = (NLIG + 3) + (NLIG + 1) + (PRECOND ? NLIG + 1 : 0);
    static constexpr int NAUX = 1;
    VecRef coef, v, pc;
    double shift, w2c;
    double *out;
    struct State {};

    __device__ __forceinline__ void load(const Geom &g, int k, long long poff,
                                         double *pre) const
    {
        const double *pc_ = plane_ptr(coef, k, g.nloc, g.plane_pts * (NLIG + 3)) + poff;
#pragma unroll
        for (int c = 0; c < NLIG + 3; ++c) pre[c] = __ldg(pc_ + c * g.plane_pts);
        const double *pv = plane_ptr(v, k, g.nloc, g.plane_pts * (NLIG + 1)) + poff;
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) pre[NLIG + 3 + c] = __ldg(pv + c * g.plane_pts);
        if (PRECOND) {
            const double *pp = plane_ptr(pc, k, g.nloc, g.plane_pts * (NLIG + 1)) + poff;
#pragma unroll
            for (int c = 0; c < NLIG + 1; ++c)
                pre[2 * NLIG + 4 + c] = __ldg(pp + c * g.plane_pts);
        }
    }
    __device__ __forceinline__ void load_aux(const Geom &, int, long long,
                                             double *) const {}
    __device__ __forceinline__ void stage(const DevPhys &P, const double *pre,
                                          double *f) const
    {
        double z[NLIG + 1];
        if (PRECOND) {
            pc_point(P, shift, w2c, pre + 2 * NLIG + 4, pre + NLIG + 3, z, NLIG);
        } else {
#pragma unroll
            for (int c = 0; c < NLIG + 1; ++c) z[c] = pre[NLIG + 3 + c];
        }
        double dG = pre[2] * z[0];
#pragma unroll
        for (int l = 0; l < NLIG; ++l) dG = fma(pre[3 + l], z[1 + l], dG);
        f[0] = z[0];
        f[1] = dG;
#pragma unroll
        for (int l = 0; l < NLIG; ++l) f[2 + l] = z[1 + l];
        f[NLIG + 2] = pre[0];
        f[NLIG + 3] = pre[1];
    }
    template <class Acc>
    __device__ __forceinline__ void emit(const DevPhys &P, const Geom &g, int ko,
                                         long long poff, const Acc &a,
                                         const double *, State &) const
    {
        constexpr int FR = NLIG + 2, FG = NLIG + 3;
        double acc = 0.0, lapG = 0.0, lapdG = 0.0;
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) {
            acc = fma(a.wsum(0, ax, P.w1[ax]), a.wsum(FG, ax, P.w1[ax]), acc);
            acc = fma(a.wsum(FR, ax, P.w1[ax]), a.wsum(1, ax, P.w1[ax]), acc);
            lapG += a.wsum(FG, ax, P.w2[ax]);
            lapdG += a.wsum(1, ax, P.w2[ax]);
        }
        const double z0 = a.c(0);
        const double Jv0 = fma(a.c(FR), lapdG, fma(z0, lapG, acc));
        const long long e = (long long)ko * (NLIG + 1) * g.plane_pts + poff;
        out[e] = fma(shift, z0, -Jv0);
#pragma unroll
        for (int l = 0; l < NLIG; ++l) {
            double lapV = 0.0;
#pragma unroll
            for (int ax = 0; ax < DIM; ++ax) lapV += a.wsum(2 + l, ax, P.w2[ax]);
            const double zl = a.c(2 + l);
            double JvU = fma(P.D[l], lapV, fma(P.s[l], z0, -P.gamma[l] * zl));
            out[e + (1 + l) * g.plane_pts] = fma(shift, zl, -JvU);
        }
    }
    __device__ __forceinline__ void finish(State &) const {}
};

// grad G and max |grad G| per axis (KSFD/ksfdsym.py:1188-1209, ksfdts.py:302-313)
template <int DIM, int NLIG>
struct VelocityOp {
    static constexpr int NF = 1;
    static constexpr int NPRE = NLIG + 1;
    static constexpr int NAUX = 1;
    VecRef u;
    double *vel;        // optional plane-SoA output with DIM fields
    double *vmax;       // optional per-axis max
    struct State {
        double vm[3];
        __device__ State() { vm[0] = vm[1] = vm[2] = 0.0; }
    };

    __device__ __forceinline__ void load(const Geom &g, int k, long long poff,
                                         double *pre) const
    {
        const double *p = plane_ptr(u, k, g.nloc, g.plane_pts * (NLIG + 1)) + poff;
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) pre[c] = __ldg(p + c * g.plane_pts);
    }
    __device__ __forceinline__ void load_aux(const Geom &, int, long long,
                                             double *) const {}
    __device__ __forceinline__ void stage(const DevPhys &P, const double *pre,
                                          double *f) const
    {
        double rho = clampv(pre[0], P.rhomin);
        double U[NLIG];
#pragma unroll
        for (int l = 0; l < NLIG; ++l) U[l] = clampv(pre[1 + l], P.Umin);
        f[0] = G_point<NLIG>(P, rho, U);
    }
    template <class Acc>
    __device__ __forceinline__ void emit(const DevPhys &P, const Geom &g, int ko,
                                         long long poff, const Acc &a,
                                         const double *, State &st) const
    {
        const long long e = (long long)ko * DIM * g.plane_pts + poff;
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) {
            double d = a.wsum(0, ax, P.w1[ax]);
            if (vel) vel[e + ax * g.plane_pts] = d;
            st.vm[ax] = fmax(st.vm[ax], fabs(d));
        }
    }
    __device__ __forceinline__ void finish(State &st) const
    {
        if (!vmax) return;
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) {
            double m = warp_max(st.vm[ax]);
            if ((threadIdx.x & 31) == 0 && m > 0.0) atomic_max_nonneg(vmax + ax, m);
        }
    }
};

// ---------------------------------------------------------------------------
// The marching skeleton
// ---------------------------------------------------------------------------
template <int DIM, int TXL, int TYL, class Op>
__global__ void __launch_bounds__(TXL *TYL)
k_march(const __grid_constant__ Geom g, const __grid_constant__ DevPhys P,
        const int RZ, Op op)
{
    constexpr int NF = Op::NF;
    constexpr int NPRE = Op::NPRE;
    constexpr int NAUX = Op::NAUX;
    constexpr int NT = TXL * TYL;
    extern __shared__ double ring[];            // [KSFD_RING][NF][NT]
    const int tid = threadIdx.x;
    const int lx = (TYL == 1) ? tid : tid % TXL;
    const int ly = (TYL == 1) ? 0 : tid / TXL;
    const int i0 = blockIdx.x * (TXL - 2 * KSFD_SW);
    const int j0 = (DIM == 3) ? blockIdx.y * (TYL - 2 * KSFD_SW) : 0;
    const int k0 = blockIdx.z * RZ;
    const int k1 = min(k0 + RZ, g.nloc);

    long long poff = wrapi(i0 - KSFD_SW + lx, g.n0);
    bool emits = lx >= KSFD_SW && lx < TXL - KSFD_SW && (i0 + lx - KSFD_SW) < g.n0;
    if (DIM == 3) {
        poff += (long long)wrapi(j0 - KSFD_SW + ly, g.n1) * g.n0;
        emits = emits && ly >= KSFD_SW && ly < TYL - KSFD_SW &&
                (j0 + ly - KSFD_SW) < g.n1;
    }
    typename Op::State st;
    double q[NF][5];
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int s = 0; s < 5; ++s) q[f][s] = 0.0;

    double pre[NPRE], aux[NAUX];
    op.load(g, k0 - KSFD_SW, poff, pre);

    int slot = 0;
    for (int kk = k0 - KSFD_SW; kk < k1 + KSFD_SW; ++kk) {
        double cur[NPRE];
#pragma unroll
        for (int c = 0; c < NPRE; ++c) cur[c] = pre[c];
        if (kk + 1 < k1 + KSFD_SW) op.load(g, kk + 1, poff, pre);
        const int ko = kk - KSFD_SW;
        const bool do_emit = emits && ko >= k0;
        if (do_emit) op.load_aux(g, ko, poff, aux);

        double f[NF];
        op.stage(P, cur, f);
#pragma unroll
        for (int c = 0; c < NF; ++c) {
#pragma unroll
            for (int s = 0; s < 4; ++s) q[c][s] = q[c][s + 1];
            q[c][4] = f[c];
            ring[(slot * NF + c) * NT + tid] = f[c];
        }
        __syncthreads();
        if (do_emit) {
            RegAcc<DIM, NF, TXL, NT> a;
            a.q = q;
            a.rc = ring + ((slot ^ 2) * NF) * NT + tid;   // slot of plane ko
            op.emit(P, g, ko, poff, a, aux, st);
        }
        slot = (slot + 1) & (KSFD_RING - 1);
    }
    op.finish(st);
}
