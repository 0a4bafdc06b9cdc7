// Preconditioned Richardson sweep fused into the marching stencil kernel.
//
// The stage systems of a ROSW step are (shift*I - J) x = b with a large shift: the
// point-block-Jacobi preconditioned operator A*M^-1 has its spectrum in a small disc
// around 1 (radius 0.09 for the benchmark problem, dt = 1e-3), so the stationary
// iteration
//        x_{k+1} = x_k + M^-1 r_k ,      r_{k+1} = r_k - A M^-1 r_k
// contracts the residual by that radius per sweep — as fast as GMRES on the same
// operator — and needs NO inner products, no basis vectors and no orthogonalisation:
// one pass of the fused A*M^-1 stencil kernel per iteration, whose emit phase also
// updates x and r and accumulates ||r_k||^2 and ||r_{k+1}||^2.  The CTA that finishes
// last sums the per-CTA partial sums in a fixed order (deterministic), adds the other
// ranks' sums over NVLink peer memory, takes the convergence decision and publishes it
// in the pinned status block the host polls (as the GMRES epilogue does); sweeps that
// were launched ahead return at once.  r_k is the TRUE residual of x_k up to rounding
// (both are advanced by the same M^-1 r_k), so the decision is taken on a true
// residual norm.  When the contraction is slow (large time steps) the host falls back
// to GMRES starting from the x reached so far (ksfd.cu: sweep_solve_impl).
//
// Replaces, like the GMRES path, the KSP solve behind TS.step() (KSFD/ksfdts.py:211;
// PETSc analogue: -ksp_type richardson -pc_type pbjacobi, true-residual norm).
#pragma once
#include "march_kernels.cuh"
#include "solver_state.cuh"

#define KSFD_SWEEP_FALLBACK 100         // reason: contraction too slow, continue with GMRES

struct SweepFin {
    // [3][cap] per-CTA partial sums: <b,b> (written by sweep 0), <r_k,r_k>, <r_{k+1},r_{k+1}>
    double *partial;
    int cap;
    int it;                     // sweep index, 0-based
    // test = 0: a sweep the solve is known to need (the host predicts the length of a solve
    // from the previous ones): no reduction, no rank sum, no decision, no epilogue at all
    int test;
    int pad_;
    double *gm;
    int *gmi;
    GmStatus *hs;
    GmOpts o;
    double slow;                // give up when ||r_{k+1}|| > slow * ||r_k||
    P2PRed pr;
    unsigned *done;             // CTA counter
};

// all threads of the CTA that finished last
__device__ __forceinline__ void sweep_finalize(const SweepFin &a, int ncta)
{
    // the three sums in one pass (this runs after everything else of the kernel: every
    // dependent step of it is exposed latency)
    __shared__ double sm3[3][32];
    __shared__ double sv[3];
    {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int b = threadIdx.x; b < ncta; b += blockDim.x) {
            t0 += a.partial[b];
            t1 += a.partial[a.cap + b];
            t2 += a.partial[2 * a.cap + b];
        }
        t0 = warp_sum(t0);
        t1 = warp_sum(t1);
        t2 = warp_sum(t2);
        if ((threadIdx.x & 31) == 0) {
            sm3[0][threadIdx.x >> 5] = t0;
            sm3[1][threadIdx.x >> 5] = t1;
            sm3[2][threadIdx.x >> 5] = t2;
        }
        __syncthreads();
        if (threadIdx.x < 3) {
            double t = 0.0;
            for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) t += sm3[threadIdx.x][i];
            sv[threadIdx.x] = t;
        }
        __syncthreads();
    }
    if (a.pr.nranks > 1) p2p_allreduce(a.pr, sv, 3, 0);
    if (threadIdx.x != 0) return;
    const double s0 = sv[0], so = sv[1], sn = sv[2];
    double *gm = a.gm;
    int *gmi = a.gmi;
    GmStatus *hs = a.hs;
    const GmOpts &o = a.o;
    const double rold = sqrt(so), rn = sqrt(sn), r0 = sqrt(s0);
    const double tol = fmax(o.rtol * r0, o.atol);
    gm[GM_RNORM0] = r0;
    gm[GM_TOL] = tol;
    const int its = a.it + 1;
    int fin = 0, reason = 0;
    if (!(rn == rn) || !(rold == rold)) { fin = 1; reason = -9; }
    else if (rn <= tol) { fin = 1; reason = r0 == 0.0 ? 3 : 2; }
    else if (o.dtol > 0.0 && rn > o.dtol * r0) { fin = 1; reason = -4; }
    else if (rn > a.slow * rold) { fin = 1; reason = KSFD_SWEEP_FALLBACK; }
    else if (its >= o.max_it) { fin = 1; reason = -3; }
    gm[GM_RNORM] = rn;
    gmi[GMI_ITS] = its;
    gmi[GMI_REASON] = reason;
    gmi[GMI_FINAL] = fin;
    gmi[GMI_CYCLE_DONE] = fin;          // later sweeps skip
    hs->rnorm = rn;
    hs->rnorm0 = r0;
    hs->its_total = its;
    if (fin) {
        // one fence: a host that sees cycle_done sees the fields above
        hs->reason = reason;
        hs->final_ = 1;
        __threadfence_system();
        hs->cycle_done = 1;
    }
    hs->iters_done = its;
}

// End of a sweep CTA, out of line (the unrolled plane loop has five exits): per-CTA partial
// sums, and in the CTA that finishes last the decision.
static __device__ __noinline__ void sweep_epilogue(const SweepFin &fin, double so_, double sn_)
{
    if (!fin.test && fin.it != 0) return;
    __shared__ double sm_[2][32];
    __shared__ int last_;
    const double a = warp_sum(so_), b = warp_sum(sn_);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    if (l == 0) {
        sm_[0][w] = a;
        sm_[1][w] = b;
    }
    __syncthreads();
    const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const int ncta = gridDim.x * gridDim.y * gridDim.z;
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int i = 0; i < nw; ++i) s += sm_[threadIdx.x][i];
        fin.partial[(1 + threadIdx.x) * fin.cap + cta] = s;
        if (fin.it == 0 && threadIdx.x == 0) fin.partial[cta] = s;      // <b,b>
    }
    if (!fin.test) return;
    if (threadIdx.x < 2) __threadfence();   // this CTA's partial sums are visible
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(fin.done, 1u);
        last_ = (t == (unsigned)ncta - 1u);
        if (last_) atomicExch(fin.done, 0u);
    }
    __syncthreads();
    if (!last_) return;
    __threadfence();                        // see every CTA's partial sums
    sweep_finalize(fin, ncta);
}

// One sweep: inputs coef, r_k (ghosted, TMA-fed like the v of J.v), pc; per output point
//   z = M^-1 r_k (staged, as in JvpOp<PRECOND>),  t = A z (the J.v stencil),
//   r_{k+1} = r_k - t  -> rout,   x_{k+1} = x_k + z  -> x (in place; first sweep: x = z)
// rsign multiplies the raw input (first sweep of a solve with right-hand side -F).
template <int DIM, int NLIG>
struct SweepOp : JvpOp<DIM, NLIG, true> {
    using Base = JvpOp<DIM, NLIG, true>;
    static constexpr int NAUX = 2 * (NLIG + 1);     // r_k and x_k of the output point
    static constexpr bool HAS_AUX = true;
    static constexpr bool NEEDS_OWNER = true;       // overlapping (clamped) tiles must not update twice
    const double *rin;          // = v.base: r_k of the owned planes
    double *x, *rout;
    double rsign;
    int first, pad_;
    SweepFin fin;
    // several ranks: the boundary planes of r_{k+1} also go to the neighbours' ghost buffers
    // (all-null descriptor: no push).  The CTAs that own boundary planes are the short
    // chunks the launcher puts first (MarchArgs::rb), so their stores and the flag are on
    // their way long before the kernel ends (cf. halo_push_publish).
    HaloPush hp;
    struct State : Base::State {
        double so, sn;
        bool own;
    };
    __device__ __forceinline__ void init_out(const MarchArgs &g, State &st, int k0, int poff) const
    {
        Base::init_out(g, st, k0, poff);
        st.so = st.sn = 0.0;
        st.own = true;
    }
    __device__ __forceinline__ void init(const MarchArgs &g, State &st, int kfirst, int k0,
                                         int poff) const
    {
        Base::init(g, st, kfirst, k0, poff);
        st.so = st.sn = 0.0;
        st.own = true;
    }
    // (TMA-fed marcher) the CTA goes on with another chunk: outputs restart, sums go on
    __device__ __forceinline__ void next_chunk(const MarchArgs &g, State &st, int k0, int poff) const
    {
        Base::init_out(g, st, k0, poff);
    }
    // (TMA-fed marcher) the chunk [k0, k1) is written: push its boundary planes
    __device__ __forceinline__ void chunk_done(const MarchArgs &g, State &st, int k0, int k1,
                                               bool mine) const
    {
        if (hp.up_lo0)
            march_push_epilogue<NLIG + 1>(hp, g, rout, k0, k1, st.e - k1 * (NLIG + 1) * g.fs, mine);
    }
    template <class Sink>
    __device__ __forceinline__ void load_aux(const MarchArgs &g, int e, int off,
                                             const Sink &sink) const
    {
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) sink.put(off + c, rin + (e + c * g.fs));
        if (!first) {
#pragma unroll
            for (int c = 0; c < NLIG + 1; ++c) sink.put(off + NLIG + 1 + c, x + (e + c * g.fs));
        }
    }
    template <class TA>
    __device__ __forceinline__ void stage(const DevPhys &P, const TA &T, const double *pre,
                                          double *f) const
    {
        double p2[Base::NPRE];
#pragma unroll
        for (int c = 0; c < Base::NPRE; ++c) p2[c] = pre[c];
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) p2[NLIG + 3 + c] *= rsign;
        Base::stage(P, T, p2, f);
    }
    template <class Acc>
    __device__ __forceinline__ void emit(const DevPhys &P, const MarchArgs &g, const Acc &a,
                                         const double *aux, State &st) const
    {
        if (!st.own) return;
        double t[NLIG + 1];
        Base::apply(P, a, t);
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) {
            const double z = a.c(c == 0 ? 0 : 1 + c);
            const double rc = rsign * aux[c];
            const double rn = rc - t[c];
            rout[st.e + c * g.fs] = rn;
            x[st.e + c * g.fs] = first ? z : aux[NLIG + 1 + c] + z;
            st.so = fma(rc, rc, st.so);
            st.sn = fma(rn, rn, st.sn);
        }
    }
    // every thread of the CTA (threads without outputs carry zeros)
    __device__ __forceinline__ void finish(const MarchArgs &g, State &st, int k0, int k1,
                                           bool mine) const
    {
        sweep_epilogue(fin, st.so, st.sn);
    }
};
