// Marching stencil kernels (the performance path, 2-D and 3-D).
//
// A CTA owns a tile of "lanes" in the non-marching axes (x in 2-D, x*y in 3-D,
// two halo lanes on every side, periodic wrap in those axes) and marches along
// the LAST axis over RZ output planes.  For every plane it
//   stage: loads the plane's points (software-prefetched one plane ahead),
//          clamps them and computes the pointwise fields the stencil needs
//          (for the residual: rho, G(rho,U) [3 log + 1 tanh], U_l) ONCE per
//          point, and stores them into a 6-slot shared-memory ring;
//   emit : after one __syncthreads, computes the 4th-order star stencil for the
//          plane two behind from the ring (x/y neighbours = neighbouring lanes,
//          z neighbours = neighbouring ring slots) and writes the output.
// So G is evaluated (L/out) * (RZ+4)/RZ times per output point instead of
// 1+4*dim times, the field vector is read once from HBM (halo re-reads hit
// L2) and no ghosted copy or G array is ever materialised.
//
// The skeleton is shared by three operators (Op policies): residual, J.v and
// velocity-max.
#pragma once
#include "device_common.cuh"
#include "naive_kernels.cuh"   // pc_point, w2c_total

struct MarchCfg {
    int LX, LY;         // lanes incl. halos (LY = 1 in 2-D)
    int RZ;             // output planes per CTA
};

// stencil access for one lane at emit time
template <int DIM, int NF>
struct RingAcc {
    const double *ring;
    int L, LX, lane;
    int s[5];           // ring slots of planes ko-2 .. ko+2
    __device__ __forceinline__ double at(int f, int slot, int ln) const
    {
        return ring[(slot * NF + f) * L + ln];
    }
    __device__ __forceinline__ double c(int f) const { return at(f, s[2], lane); }
    // weighted 5-point sum along axis ax with weights w[5]
    __device__ __forceinline__ double wsum(int f, int ax, const double *w) const
    {
        double r;
        if (ax == DIM - 1) {
            r = w[0] * at(f, s[0], lane);
            r = fma(w[1], at(f, s[1], lane), r);
            r = fma(w[2], at(f, s[2], lane), r);
            r = fma(w[3], at(f, s[3], lane), r);
            r = fma(w[4], at(f, s[4], lane), r);
        } else {
            const int st = (ax == 0) ? 1 : LX;
            const double *b = ring + (s[2] * NF + f) * L + lane;
            r = w[0] * b[-2 * st];
            r = fma(w[1], b[-st], r);
            r = fma(w[2], b[0], r);
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
    VecRef u;
    const double *udot, *src;
    double *out;

    __device__ __forceinline__ void load(const Geom &g, int k, long long poff,
                                         double *pre) const
    {
        const double *p =
            plane_ptr(u, k, g.nloc, g.plane_pts * (NLIG + 1)) + poff * (NLIG + 1);
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) pre[c] = __ldg(p + c);
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
    struct State {};
    template <class Acc>
    __device__ __forceinline__ void emit(const DevPhys &P, const Geom &g, int ko,
                                         long long poff, const Acc &a,
                                         State &) const
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
        const long long e = ((long long)ko * g.plane_pts + poff) * (NLIG + 1);
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) {
            double v = f[c];
            if (src) v += __ldg(src + e + c);
            out[e + c] = udot ? __ldg(udot + e + c) - v : v;
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
    VecRef coef, v, pc;
    double shift, w2c;
    double *out;

    __device__ __forceinline__ void load(const Geom &g, int k, long long poff,
                                         double *pre) const
    {
        const double *pc_ =
            plane_ptr(coef, k, g.nloc, g.plane_pts * (NLIG + 3)) + poff * (NLIG + 3);
#pragma unroll
        for (int c = 0; c < NLIG + 3; ++c) pre[c] = __ldg(pc_ + c);
        const double *pv =
            plane_ptr(v, k, g.nloc, g.plane_pts * (NLIG + 1)) + poff * (NLIG + 1);
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) pre[NLIG + 3 + c] = __ldg(pv + c);
        if (PRECOND) {
            const double *pp = plane_ptr(pc, k, g.nloc, g.plane_pts * (NLIG + 1)) +
                               poff * (NLIG + 1);
#pragma unroll
            for (int c = 0; c < NLIG + 1; ++c)
                pre[2 * NLIG + 4 + c] = __ldg(pp + c);
        }
    }
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
    struct State {};
    template <class Acc>
    __device__ __forceinline__ void emit(const DevPhys &P, const Geom &g, int ko,
                                         long long poff, const Acc &a,
                                         State &) const
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
        const long long e = ((long long)ko * g.plane_pts + poff) * (NLIG + 1);
        out[e] = fma(shift, z0, -Jv0);
#pragma unroll
        for (int l = 0; l < NLIG; ++l) {
            double lapV = 0.0;
#pragma unroll
            for (int ax = 0; ax < DIM; ++ax) lapV += a.wsum(2 + l, ax, P.w2[ax]);
            const double zl = a.c(2 + l);
            double JvU = fma(P.D[l], lapV, fma(P.s[l], z0, -P.gamma[l] * zl));
            out[e + 1 + l] = fma(shift, zl, -JvU);
        }
    }
    __device__ __forceinline__ void finish(State &) const {}
};

// max |grad G| per axis (KSFD/ksfdsym.py:1188-1209 + ksfdts.py:302-313)
template <int DIM, int NLIG>
struct VelocityOp {
    static constexpr int NF = 1;
    static constexpr int NPRE = NLIG + 1;
    VecRef u;
    double *vel;        // optional (dim, pts) output
    double *vmax;       // optional per-axis max
    struct State {
        double vm[3];
        __device__ State() { vm[0] = vm[1] = vm[2] = 0.0; }
    };

    __device__ __forceinline__ void load(const Geom &g, int k, long long poff,
                                         double *pre) const
    {
        const double *p =
            plane_ptr(u, k, g.nloc, g.plane_pts * (NLIG + 1)) + poff * (NLIG + 1);
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) pre[c] = __ldg(p + c);
    }
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
                                         State &st) const
    {
        const long long p = (long long)ko * g.plane_pts + poff;
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) {
            double d = a.wsum(0, ax, P.w1[ax]);
            if (vel) vel[p * DIM + ax] = d;
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
template <int DIM, int MAXLPT, class Op>
__global__ void k_march(const __grid_constant__ Geom g,
                        const __grid_constant__ DevPhys P,
                        const __grid_constant__ MarchCfg cfg, Op op)
{
    constexpr int NF = Op::NF;
    constexpr int NPRE = Op::NPRE;
    extern __shared__ double ring[];            // [RING][NF][L]
    const int LX = cfg.LX, LY = (DIM == 3) ? cfg.LY : 1;
    const int L = LX * LY;
    const int NT = blockDim.x;
    const int i0 = blockIdx.x * (LX - 2 * KSFD_SW);
    const int j0 = (DIM == 3) ? blockIdx.y * (LY - 2 * KSFD_SW) : 0;
    const int k0 = blockIdx.z * cfg.RZ;
    const int k1 = min(k0 + cfg.RZ, g.nloc);

    // per-lane constants
    long long poff[MAXLPT];     // point offset inside a plane (wrapped)
    bool live[MAXLPT], emits[MAXLPT];
#pragma unroll
    for (int m = 0; m < MAXLPT; ++m) {
        const int lane = threadIdx.x + m * NT;
        live[m] = lane < L;
        int lx = lane % LX, ly = lane / LX;
        int gi = wrapi(i0 - KSFD_SW + lx, g.n0);
        bool e = live[m] && lx >= KSFD_SW && lx < LX - KSFD_SW &&
                 (i0 + lx - KSFD_SW) < g.n0;
        long long po = gi;
        if (DIM == 3) {
            int gj = wrapi(j0 - KSFD_SW + ly, g.n1);
            po += (long long)gj * g.n0;
            e = e && ly >= KSFD_SW && ly < LY - KSFD_SW &&
                (j0 + ly - KSFD_SW) < g.n1;
        }
        poff[m] = po;
        emits[m] = e;
    }
    typename Op::State st;

    double pre[MAXLPT][NPRE];
#pragma unroll
    for (int m = 0; m < MAXLPT; ++m)
        if (live[m]) op.load(g, k0 - KSFD_SW, poff[m], pre[m]);

    int slot = 0;
    for (int kk = k0 - KSFD_SW; kk < k1 + KSFD_SW; ++kk) {
        double cur[MAXLPT][NPRE];
#pragma unroll
        for (int m = 0; m < MAXLPT; ++m)
#pragma unroll
            for (int c = 0; c < NPRE; ++c) cur[m][c] = pre[m][c];
        if (kk + 1 < k1 + KSFD_SW) {
#pragma unroll
            for (int m = 0; m < MAXLPT; ++m)
                if (live[m]) op.load(g, kk + 1, poff[m], pre[m]);
        }
#pragma unroll
        for (int m = 0; m < MAXLPT; ++m) {
            if (live[m]) {
                double f[NF];
                op.stage(P, cur[m], f);
                const int lane = threadIdx.x + m * NT;
#pragma unroll
                for (int q = 0; q < NF; ++q) ring[(slot * NF + q) * L + lane] = f[q];
            }
        }
        __syncthreads();
        const int ko = kk - KSFD_SW;
        if (ko >= k0) {
#pragma unroll
            for (int m = 0; m < MAXLPT; ++m) {
                if (emits[m]) {
                    RingAcc<DIM, NF> a;
                    a.ring = ring;
                    a.L = L;
                    a.LX = LX;
                    a.lane = threadIdx.x + m * NT;
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        int sq = slot + 2 + q;          // == slot - 4 + q (mod 6)
                        a.s[q] = sq >= KSFD_RING ? sq - KSFD_RING : sq;
                    }
                    op.emit(P, g, ko, poff[m], a, st);
                }
            }
        }
        slot = (slot + 1 == KSFD_RING) ? 0 : slot + 1;
    }
    op.finish(st);
}
