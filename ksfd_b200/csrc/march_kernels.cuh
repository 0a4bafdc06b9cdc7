// Marching stencil kernels (the performance path, 2-D and 3-D).
//
// A CTA owns a tile of up to TX x TY output points in the non-marching axes
// (x in 2-D, x*y in 3-D) plus the two-deep star halo around it (periodic
// wrap), one LANE per point, one thread per lane, and marches along the LAST
// axis over `rz` output planes.  Per plane kk:
//   stage: the plane's points (plane-SoA => coalesced loads, prefetched one
//          plane ahead into registers) are clamped and turned into the
//          pointwise fields the stencil needs — for the residual: rho,
//          G(rho,U) [3 log + 1 exp, table driven, evaluated ONCE per lane and
//          plane], U_l — and pushed into a 5-deep REGISTER queue per lane (the
//          marching-axis stencil costs no memory traffic);
//   share: the queue's CENTRE plane (kk-2) is written to one of two
//          shared-memory slots; one __syncthreads per plane (two slots are
//          enough: a slot is rewritten two barriers after it was read);
//   emit : interior lanes form the 4th-order star stencil of plane kk-2 from
//          the queue (marching axis) and from neighbouring lanes of the slot
//          (cross axes: immediate-offset LDS, strides are compile-time) and
//          store the outputs coalesced.
// 3-D lane packing: threads [0, TX*TY) are the interior lanes, the 4*TX + 4*TY
// halo lanes follow, so the emit phase runs on fully populated warps and halo
// lanes cost staging work only.
//
// Constants: on sm_100 fp64 instructions take constants from the 63 uniform
// registers only, so constant count is a first-order cost.  The stencils are
// therefore evaluated in their symmetric form
//     d/dx   = c1 (8 (f[+1]-f[-1]) - (f[+2]-f[-2]))
//     d2/dx2 = c2 (16 (f[+1]+f[-1]) - (f[+2]+f[-2]) - 30 f[0])
// (8, 16, 30 are instruction immediates; one scale per axis instead of ten
// weights), and log/exp are table driven (fastmath.cuh).
//
// The field vector is read once from HBM (halo re-reads hit L2), G is evaluated
// lanes/outputs * (rz+4)/rz times per output point instead of 1+4*dim times,
// and no ghosted copy or G array is ever materialised.
//
// The skeleton is shared by three operators (Op policies): residual, J.v
// (optionally fused with the block-Jacobi solve), velocity(-max).
#pragma once
#include "device_common.cuh"
#include "fastmath.cuh"
#include "halo_push.cuh"

// uniform launch parameters (32-bit: a rank-local vector has < 2^31 elements,
// checked on the host)
struct MarchArgs {
    int n0, n1, nloc;       // extents in x, y (1 in 2-D); owned planes
    int fs;                 // field stride = points per plane
    int ox, oy;             // outputs per tile in x, y (balanced pitch <= TX, TY)
    int rz;                 // output planes per CTA
    // rb > 0 (Richardson sweeps on several ranks, sweep_op.cuh; TMA-fed marcher only): the
    // boundary planes are owned by the FIRST CTA of every column, which marches two short
    // chunks one after the other — pass 0 = the top planes [nloc-rb, nloc), pass 1 = the
    // bottom planes [0, rb) — and pushes each to the neighbour as soon as it is done: the
    // top planes (which the upper neighbour's next sweep reads FIRST) are on their way a
    // third into the kernel.  Chunks 1.. = rz planes each of [rb, nloc-rb).
    int rb;
};

// planes [k0, k1) of chunk z (rb mode: pass 0 / 1 of chunk 0)
__device__ __forceinline__ void march_chunk(const MarchArgs &g, int z, int &k0, int &k1, int pass = 0)
{
    if (g.rb > 0) {
        if (z == 0) {
            k0 = pass == 0 ? g.nloc - g.rb : 0;
            k1 = k0 + g.rb;
        } else {
            k0 = g.rb + (z - 1) * g.rz;
            k1 = min(k0 + g.rz, g.nloc - g.rb);
        }
    } else {
        k0 = z * g.rz;
        k1 = min(k0 + g.rz, g.nloc);
    }
}

// End of a CTA of a marching kernel that also PUSHES the boundary planes of its output
// (several ranks, NVLink peer memory): the CTAs whose chunk [k0, k1) holds boundary planes
// copy them from the output they just wrote (every thread re-reads its own stores) into the
// neighbours' ghost buffers and arrive on the block counter; the last of them publishes
// the exchange (halo_push_publish).  The exchange counter is read HERE: it cannot advance
// before this CTA has arrived.  Out of line, every thread of the CTA calls it.
// nfs = doubles per plane of the output, poff = this thread's column, mine = it has one.
template <int NC>
static __device__ __noinline__ void march_push_epilogue(const HaloPush &hp, const MarchArgs &g,
                                                        const double *out, int k0, int k1,
                                                        int poff, bool mine)
{
    if (!halo_push_on(hp)) return;
    if (!(k0 < KSFD_SW || k1 > g.nloc - KSFD_SW)) return;       // uniform over the CTA
    unsigned long long q;
    const long long sh = halo_push_shift(hp, q);
    if (mine) {
        const int nfs = NC * g.fs;
        for (int k = k0; k < k1; ++k) {
            if (k >= KSFD_SW && k < g.nloc - KSFD_SW) continue;
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const long long e = (long long)k * nfs + c * g.fs + poff;
                halo_push1(hp, sh, e, out[e]);
            }
        }
    }
    if (g.rb > 0) {
        // one boundary CTA per column, two passes: the top planes' flag goes out after pass 0,
        // the bottom planes' flag and the exchange counter after pass 1
        halo_push_publish_dir(hp, q, gridDim.x * gridDim.y, k0 == 0 ? 1 : 0);
        return;
    }
    unsigned nbz = 0;           // chunks that hold boundary planes
    for (int z = 0; z < (int)gridDim.z; ++z) {
        int a, b;
        march_chunk(g, z, a, b);
        nbz += (a < KSFD_SW || b > g.nloc - KSFD_SW) ? 1u : 0u;
    }
    halo_push_publish(hp, q, nbz * gridDim.x * gridDim.y);
}

template <int DIM, int TX, int TY>
struct TileT {
    static constexpr int NL = (DIM == 2) ? TX + 4 : TX * TY + 4 * TX + 4 * TY;   // lanes
    static constexpr int NT = (NL + 31) / 32 * 32;                                // threads
    static constexpr int PX = TX + 4;                               // smem row pitch
    static constexpr int SP = (DIM == 2) ? TX + 4 : (TX + 4) * (TY + 4);
    static constexpr int SY = (DIM == 2) ? 0 : PX;
};

// log / exp tables (fastmath.cuh), staged to shared memory by each CTA
static __device__ const unsigned long long g_log_tab[256] = KSFD_LOG_TAB_INIT;
static __device__ const unsigned long long g_exp_tab[64] = KSFD_EXP_TAB_INIT;
#define KSFD_TAB_DOUBLES 320

// All shared memory is addressed through this symbol with integer indices (a
// generic pointer into shared memory would make every access recompute the
// shared window base).  Layout: [2][NF][SP] ring, then the tables.
extern __shared__ __align__(128) double ksfd_smem[];

// table accessor over the CTA's shared-memory copy; OFF = index of the tables
template <int OFF>
struct SmemTabs {
    __device__ __forceinline__ void log_pair(int i, double &invc, double &logc) const
    {
        const double2 v = *reinterpret_cast<const double2 *>(&ksfd_smem[OFF + 2 * i]);
        invc = v.x;
        logc = v.y;
    }
    __device__ __forceinline__ double exp2j(int j) const { return ksfd_smem[OFF + 256 + j]; }
};

// pointer to plane k of a vector whose planes hold `ps` doubles
__device__ __forceinline__ const double *plane_of(const VecRef &v, int k, int nloc, int ps)
{
    if (k >= 0 && k < nloc) return v.base + k * (long long)ps;
    // ghost plane: wait (every thread of the CTA, they all load from it) until the
    // neighbour's planes of the current exchange have landed
    const unsigned long long q =
        v.par ? *reinterpret_cast<const volatile unsigned long long *>(v.par) : 0ull;
    const long long shift = (long long)(q & 1ull) * v.pstride;
    const volatile unsigned long long *f = k < 0 ? v.flag_lo : v.flag_hi;
    if (f) halo_flag_wait(f, q, v.err, v.dead);
    if (k < 0) return v.lo + shift + (k + KSFD_SW) * (long long)ps;
    return v.hi + shift + (k - nloc) * (long long)ps;
}

// input cursor: per-thread pointer to this lane's point in the next plane to
// load; planes are contiguous except at the two ghost boundaries
struct InCursor {
    const double *p;
    __device__ __forceinline__ void init(const VecRef &v, int k, int nloc, int ps, int poff)
    {
        p = plane_of(v, k, nloc, ps) + poff;
    }
    // after plane k was loaded
    __device__ __forceinline__ void next(const VecRef &v, int k, int nloc, int ps, int poff)
    {
        const int kn = k + 1;
        if (kn == 0 || kn == nloc)              // uniform, at most twice per CTA
            p = plane_of(v, kn, nloc, ps) + poff;
        else
            p += ps;
    }
};

// The loads of a plane go straight into registers.  Plain coherent loads, not the
// non-coherent read-only path: with several ranks the ghost planes are written by the
// neighbouring GPU WHILE this kernel runs (it waits for them in plane_of), which ld.global.nc
// does not allow; ghost lines are first touched after the wait, so the L1 holds no stale copy.
struct RegSink {
    double *r;
    __device__ __forceinline__ void put(int c, const double *p) const
    {
        double v;
        asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
        r[c] = v;
    }
};

// stencil access of one interior lane at emit time.  PH = phase of the
// unrolled plane loop: plane kk (newest) sits in q[.][PH], plane kk-4+s in
// q[.][(PH+1+s)%5], the centre plane kk-2 in q[.][(PH+3)%5].
template <int DIM, int NF, int SP, int SY, int PH>
struct LaneAcc {
    const double (&q)[NF][5];
    const int ri;                // index of slot[field 0][this lane] in ksfd_smem
    __device__ __forceinline__ LaneAcc(const double (&q_)[NF][5], int ri_)
        : q(q_), ri(ri_) {}
    __device__ __forceinline__ double c(int f) const { return q[f][(PH + 3) % 5]; }
    // the four neighbours of field f along axis ax
    __device__ __forceinline__ void nb(int f, int ax, double &m2, double &m1, double &p1,
                                       double &p2) const
    {
        if (ax == DIM - 1) {
            m2 = q[f][(PH + 1) % 5];
            m1 = q[f][(PH + 2) % 5];
            p1 = q[f][(PH + 4) % 5];
            p2 = q[f][PH];
        } else {
            const int st = (ax == 0) ? 1 : SY;
            const int b = ri + f * SP;
            m2 = ksfd_smem[b - 2 * st];
            m1 = ksfd_smem[b - st];
            p1 = ksfd_smem[b + st];
            p2 = ksfd_smem[b + 2 * st];
        }
    }
    // 12h * d/dx:  8 (f[+1]-f[-1]) - (f[+2]-f[-2])
    __device__ __forceinline__ double D1(int f, int ax) const
    {
        double m2, m1, p1, p2;
        nb(f, ax, m2, m1, p1, p2);
        return fma(8.0, p1 - m1, m2 - p2);
    }
    // 12h^2 * d2/dx2:  16 (f[+1]+f[-1]) - (f[+2]+f[-2]) - 30 f[0]
    __device__ __forceinline__ double D2(int f, int ax) const
    {
        double m2, m1, p1, p2;
        nb(f, ax, m2, m1, p1, p2);
        return fma(-30.0, c(f), fma(16.0, p1 + m1, -(p2 + m2)));
    }
    // both at once (shares the neighbour loads)
    __device__ __forceinline__ void D12(int f, int ax, double &d1, double &d2) const
    {
        double m2, m1, p1, p2;
        nb(f, ax, m2, m1, p1, p2);
        d1 = fma(8.0, p1 - m1, m2 - p2);
        d2 = fma(-30.0, c(f), fma(16.0, p1 + m1, -(p2 + m2)));
    }
};

// the log part of G for arguments outside the fast domain (never taken for
// physical states; kept out of line, arguments by value)
static __device__ __noinline__ double G_logs_slow(double s2, double rho, int ng, double b0,
                                                  double a0, double b1, double a1, double b2,
                                                  double a2, double b3, double a3)
{
    double G = s2 * log(rho);
    if (ng > 0) G = fma(-b0, log(a0), G);
    if (ng > 1) G = fma(-b1, log(a1), G);
    if (ng > 2) G = fma(-b2, log(a2), G);
    if (ng > 3) G = fma(-b3, log(a3), G);
    return G;
}

// Pointwise free energy G(rho, U) = V + s2*log(rho) with the table-driven
// log / logistic (KSFD/ksfdsym.py:983-990; KSFD/ksfdligand.py:527-547;
// ksfdsoln.py:147-161)
template <int NLIG, class TA>
__device__ __forceinline__ double G_fast(const DevPhys &P, const TA &T, double rho,
                                         const double *U)
{
    // all NLIG group slots are evaluated without tests: slots beyond ngroups
    // hold alpha = 1, beta = 0, W = 0 (set on the host), i.e. contribute
    // -0*log(1); straight-line code lets the independent log chains interleave
    double a[NLIG];
    bool ok = ksfd_log_domain(rho);
#pragma unroll
    for (int g = 0; g < NLIG; ++g) {
        double t = P.alpha[g];
#pragma unroll
        for (int l = 0; l < NLIG; ++l) t = fma(P.Wgl[g][l], U[l], t);
        a[g] = t;
        ok = ok && ksfd_log_domain(t);
    }
    double G;
    if (ok) {
        G = P.s2 * ksfd_log_core(rho, P.mk, T);
#pragma unroll
        for (int g = 0; g < NLIG; ++g) G = fma(-P.beta[g], ksfd_log_core(a[g], P.mk, T), G);
    } else {
        G = G_logs_slow(P.s2, rho, NLIG, P.beta[0], a[0], P.beta[1], a[NLIG > 1 ? 1 : 0],
                        P.beta[2], a[NLIG > 2 ? 2 : 0], P.beta[3], a[NLIG > 3 ? 3 : 0]);
    }
    // capscale*(tanh(x)+1) = 2*capscale/(1+exp(-2x)),  x = (rho-rhomax)/cushion
    double cap = P.capscale2 * ksfd_logistic(fma(P.ycap1, rho, P.ycap0), P.mk, T);
    if (P.cap_type == 1) cap *= rho * P.inv_rhomax;
    return G + cap;
}

// z = M^{-1} r for one point of the point-block Jacobi preconditioner.
//   block = [ a  b_l ; c_l  d_l ],  c_l = -s_l,  d_l = shift + gamma_l - D_l*w2c
//   b_l = -(rho*w2c) * dG/dU_l   (the w1-centre term of the exact block, zero up
//   to the last-bit asymmetry of the reference weights, is left to the Krylov
//   iteration);  pcinv = 1/(a - sum_l b_l c_l / d_l)  is the only stored field.
template <int NLIG>
__device__ __forceinline__ void pc_solve(const DevPhys &P, const double *invd,
                                         double rho, const double *gU, double pcinv,
                                         const double *r, double *z)
{
    const double fac = rho * P.w2c;
    double t = r[0];
#pragma unroll
    for (int l = 0; l < NLIG; ++l) t = fma(fac * gU[l] * invd[l], r[1 + l], t);
    const double zr = pcinv * t;
    z[0] = zr;
#pragma unroll
    for (int l = 0; l < NLIG; ++l) z[1 + l] = fma(P.s[l], zr, r[1 + l]) * invd[l];
}

// ---------------------------------------------------------------------------
// Operator policies.  Each owns its per-thread pointers (State): inputs advance
// plane by plane (InCursor), outputs advance by one plane per emit.
// ---------------------------------------------------------------------------
// Residual: F = udot - (f(u)+src)  |  f(u)+src
// (KSFD/ksfdsym.py:902-940 + KSFD/ksfdts.py:591-592 fused)
// FIXED = true: udot given, no source (the implicit-step case; no runtime tests)
template <int DIM, int NLIG, bool FIXED>
struct ResidualOp {
    static constexpr int NF = NLIG + 2;        // rho, G, U_l
    static constexpr int NPRE = NLIG + 1;
    static constexpr int NAUX = NLIG + 1;      // udot of the output plane
    static constexpr bool HAS_AUX = true;
    static constexpr bool TABS = true;
    static constexpr bool JACOBIAN = false;     // runs on the current physics (ctx.P)
    static constexpr bool NEEDS_OWNER = false;
    // input vectors of the TMA-fed marcher (tma_march.cuh): u
    static constexpr int NIN = 1;
    __host__ __device__ static constexpr int nc(int) { return NLIG + 1; }
    __host__ __device__ static constexpr int coff(int) { return 0; }
    VecRef u;
    const double *udot, *src;
    double *out;
    // several ranks: the boundary planes of the result also go to the neighbours' ghost
    // buffers (all-null: no push) — the right-hand side of the first Richardson sweep
    HaloPush hp;
    struct State {
        InCursor in;
        int e;                                 // element index of the next output
    };
    __device__ static constexpr int out_fields(int) { return NLIG + 1; }
    __device__ __forceinline__ void init_out(const MarchArgs &g, State &st, int k0, int poff) const
    {
        st.e = k0 * (NLIG + 1) * g.fs + poff;
    }

    __device__ __forceinline__ void init(const MarchArgs &g, State &st, int kfirst, int k0,
                                         int poff) const
    {
        st.in.init(u, kfirst, g.nloc, g.fs * (NLIG + 1), poff);
        st.e = k0 * (NLIG + 1) * g.fs + poff;
    }
    template <class Sink>
    __device__ __forceinline__ void load(const MarchArgs &g, State &st, int k, int poff,
                                         const Sink &sink) const
    {
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) sink.put(c, st.in.p + c * g.fs);
        st.in.next(u, k, g.nloc, g.fs * (NLIG + 1), poff);
    }
    // udot of output plane ko (element index e of this lane), fields [off, off+NAUX)
    template <class Sink>
    __device__ __forceinline__ void load_aux(const MarchArgs &g, int e, int off,
                                             const Sink &sink) const
    {
        if (FIXED || udot) {
#pragma unroll
            for (int c = 0; c < NLIG + 1; ++c) sink.put(off + c, udot + (e + c * g.fs));
        }
    }
    template <class TA>
    __device__ __forceinline__ void stage(const DevPhys &P, const TA &T, const double *pre,
                                          double *f) const
    {
        const double rho = clampv(pre[0], P.rhomin);
        double U[NLIG];
#pragma unroll
        for (int l = 0; l < NLIG; ++l) U[l] = clampv(pre[1 + l], P.Umin);
        f[0] = rho;
        f[1] = G_fast<NLIG>(P, T, rho, U);
#pragma unroll
        for (int l = 0; l < NLIG; ++l) f[2 + l] = U[l];
    }
    template <class Acc>
    __device__ __forceinline__ void emit(const DevPhys &P, const MarchArgs &g, const Acc &a,
                                         const double *aux, State &st) const
    {
        // f_rho = grad(rho).grad(G) + rho*lap(G)
        const double rho0 = a.c(0);
        double acc = 0.0, lap = 0.0;
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) {
            double d1G, d2G;
            a.D12(1, ax, d1G, d2G);
            acc = fma(a.D1(0, ax) * P.c1sq[ax], d1G, acc);
            lap = fma(P.c2[ax], d2G, lap);
        }
        double f[NLIG + 1];
        f[0] = fma(rho0, lap, acc);
#pragma unroll
        for (int l = 0; l < NLIG; ++l) {
            double lapU = 0.0;
#pragma unroll
            for (int ax = 0; ax < DIM; ++ax) lapU = fma(P.c2[ax], a.D2(2 + l, ax), lapU);
            f[1 + l] =
                fma(P.D[l], lapU, fma(P.s[l], rho0, -P.gamma[l] * a.c(2 + l)));
        }
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) {
            double v = f[c];
            if (FIXED) {
                v = aux[c] - v;
            } else {
                if (src) v += __ldg(src + (st.e + c * g.fs));
                if (udot) v = aux[c] - v;
            }
            out[st.e + c * g.fs] = v;
        }
    }
    // once per emitted plane, by every thread
    __device__ __forceinline__ void advance_out(const MarchArgs &g, State &st) const
    {
        st.e += (NLIG + 1) * g.fs;
    }
    // every thread of the CTA; [k0, k1) = its planes, mine = this thread emitted
    __device__ __forceinline__ void finish(const MarchArgs &g, State &st, int k0, int k1,
                                           bool mine) const
    {
        if (hp.up_lo0)
            march_push_epilogue<NLIG + 1>(hp, g, out, k0, k1, st.e - k1 * (NLIG + 1) * g.fs, mine);
    }
};

// J.v: out = (shift*I - J(u_lin)) * z,  z = v or M^{-1} v
// (replaces the assembled matrix of KSFD/ksfdsym.py:814-886)
// coef fields per point: rho, G, dG/drho, dG/dU_l  (NLIG+3)
template <int DIM, int NLIG, bool PRECOND>
struct JvpOp {
    static constexpr int NF = NLIG + 4;   // z_rho, dG, z_U.., rho, G
    static constexpr int NPRE = (NLIG + 3) + (NLIG + 1) + (PRECOND ? 1 : 0);
    static constexpr int NAUX = 1;
    static constexpr bool HAS_AUX = false;
    static constexpr bool TABS = false;
    static constexpr bool JACOBIAN = true;      // runs on the physics of the linearisation (ctx.Pjac)
    static constexpr bool NEEDS_OWNER = false;   // (sweep_op.cuh) outputs updated in place: one owner per point
    // input vectors of the TMA-fed marcher (tma_march.cuh): coef, v, pc
    static constexpr int NIN = PRECOND ? 3 : 2;
    __host__ __device__ static constexpr int nc(int i) { return i == 0 ? NLIG + 3 : i == 1 ? NLIG + 1 : 1; }
    __host__ __device__ static constexpr int coff(int i) { return i == 0 ? 0 : i == 1 ? NLIG + 3 : 2 * NLIG + 4; }
    VecRef coef, v, pc;
    double shift;
    double invd[NLIG];
    double *out;
    struct State {
        InCursor ic, iv, ip;
        int e;
    };
    __device__ static constexpr int out_fields(int) { return NLIG + 1; }
    __device__ __forceinline__ void init_out(const MarchArgs &g, State &st, int k0, int poff) const
    {
        st.e = k0 * (NLIG + 1) * g.fs + poff;
    }

    __device__ __forceinline__ void init(const MarchArgs &g, State &st, int kfirst, int k0,
                                         int poff) const
    {
        st.ic.init(coef, kfirst, g.nloc, g.fs * (NLIG + 3), poff);
        st.iv.init(v, kfirst, g.nloc, g.fs * (NLIG + 1), poff);
        if (PRECOND) st.ip.init(pc, kfirst, g.nloc, g.fs, poff);
        st.e = k0 * (NLIG + 1) * g.fs + poff;
    }
    template <class Sink>
    __device__ __forceinline__ void load(const MarchArgs &g, State &st, int k, int poff,
                                         const Sink &sink) const
    {
#pragma unroll
        for (int c = 0; c < NLIG + 3; ++c) sink.put(c, st.ic.p + c * g.fs);
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) sink.put(NLIG + 3 + c, st.iv.p + c * g.fs);
        if (PRECOND) sink.put(2 * NLIG + 4, st.ip.p);
        st.ic.next(coef, k, g.nloc, g.fs * (NLIG + 3), poff);
        st.iv.next(v, k, g.nloc, g.fs * (NLIG + 1), poff);
        if (PRECOND) st.ip.next(pc, k, g.nloc, g.fs, poff);
    }
    template <class Sink>
    __device__ __forceinline__ void load_aux(const MarchArgs &, int, int, const Sink &) const {}
    template <class TA>
    __device__ __forceinline__ void stage(const DevPhys &P, const TA &, const double *pre,
                                          double *f) const
    {
        double z[NLIG + 1];
        if (PRECOND) {
            pc_solve<NLIG>(P, invd, pre[0], pre + 3, pre[2 * NLIG + 4], pre + NLIG + 3, z);
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
    // t = (shift*I - J) z at one output point (all NLIG+1 components)
    template <class Acc>
    __device__ __forceinline__ void apply(const DevPhys &P, const Acc &a, double *t) const
    {
        // (J z)_rho = grad(z0).grad(G) + grad(rho).grad(dG) + z0*lap(G) + rho*lap(dG)
        constexpr int FR = NLIG + 2, FG = NLIG + 3;
        double acc = 0.0, lapG = 0.0, lapdG = 0.0;
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) {
            double d1G, d2G, d1dG, d2dG;
            a.D12(FG, ax, d1G, d2G);
            a.D12(1, ax, d1dG, d2dG);
            double t_ = a.D1(0, ax) * d1G;
            t_ = fma(a.D1(FR, ax), d1dG, t_);
            acc = fma(P.c1sq[ax], t_, acc);
            lapG = fma(P.c2[ax], d2G, lapG);
            lapdG = fma(P.c2[ax], d2dG, lapdG);
        }
        const double z0 = a.c(0);
        const double Jv0 = fma(a.c(FR), lapdG, fma(z0, lapG, acc));
        t[0] = fma(shift, z0, -Jv0);
#pragma unroll
        for (int l = 0; l < NLIG; ++l) {
            double lapV = 0.0;
#pragma unroll
            for (int ax = 0; ax < DIM; ++ax) lapV = fma(P.c2[ax], a.D2(2 + l, ax), lapV);
            const double zl = a.c(2 + l);
            const double JvU = fma(P.D[l], lapV, fma(P.s[l], z0, -P.gamma[l] * zl));
            t[1 + l] = fma(shift, zl, -JvU);
        }
    }
    template <class Acc>
    __device__ __forceinline__ void emit(const DevPhys &P, const MarchArgs &g, const Acc &a,
                                         const double *, State &st) const
    {
        double t[NLIG + 1];
        apply(P, a, t);
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) out[st.e + c * g.fs] = t[c];
    }
    __device__ __forceinline__ void advance_out(const MarchArgs &g, State &st) const
    {
        st.e += (NLIG + 1) * g.fs;
    }
    __device__ __forceinline__ void finish(const MarchArgs &, State &, int, int, bool) const {}
};

// grad G and max |grad G| per axis (KSFD/ksfdsym.py:1188-1209, ksfdts.py:302-313)
template <int DIM, int NLIG>
struct VelocityOp {
    static constexpr int NF = 1;
    static constexpr int NPRE = NLIG + 1;
    static constexpr int NAUX = 1;
    static constexpr bool HAS_AUX = false;
    static constexpr bool TABS = true;
    static constexpr bool JACOBIAN = false;
    static constexpr bool NEEDS_OWNER = false;
    static constexpr int NIN = 1;
    __host__ __device__ static constexpr int nc(int) { return NLIG + 1; }
    __host__ __device__ static constexpr int coff(int) { return 0; }
    VecRef u;
    double *vel;        // optional plane-SoA output with DIM fields
    double *vmax;       // optional per-axis max
    struct State {
        InCursor in;
        int e;
        double vm[3];
    };
    __device__ static constexpr int out_fields(int dim) { return dim; }
    __device__ __forceinline__ void init_out(const MarchArgs &g, State &st, int k0, int poff) const
    {
        st.e = k0 * DIM * g.fs + poff;
        st.vm[0] = st.vm[1] = st.vm[2] = 0.0;
    }

    __device__ __forceinline__ void init(const MarchArgs &g, State &st, int kfirst, int k0,
                                         int poff) const
    {
        st.in.init(u, kfirst, g.nloc, g.fs * (NLIG + 1), poff);
        st.e = k0 * DIM * g.fs + poff;
        st.vm[0] = st.vm[1] = st.vm[2] = 0.0;
    }
    template <class Sink>
    __device__ __forceinline__ void load(const MarchArgs &g, State &st, int k, int poff,
                                         const Sink &sink) const
    {
#pragma unroll
        for (int c = 0; c < NLIG + 1; ++c) sink.put(c, st.in.p + c * g.fs);
        st.in.next(u, k, g.nloc, g.fs * (NLIG + 1), poff);
    }
    template <class Sink>
    __device__ __forceinline__ void load_aux(const MarchArgs &, int, int, const Sink &) const {}
    template <class TA>
    __device__ __forceinline__ void stage(const DevPhys &P, const TA &T, const double *pre,
                                          double *f) const
    {
        const double rho = clampv(pre[0], P.rhomin);
        double U[NLIG];
#pragma unroll
        for (int l = 0; l < NLIG; ++l) U[l] = clampv(pre[1 + l], P.Umin);
        f[0] = G_fast<NLIG>(P, T, rho, U);
    }
    template <class Acc>
    __device__ __forceinline__ void emit(const DevPhys &P, const MarchArgs &g, const Acc &a,
                                         const double *, State &st) const
    {
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) {
            const double d = P.c1[ax] * a.D1(0, ax);
            if (vel) vel[st.e + ax * g.fs] = d;
            st.vm[ax] = fmax(st.vm[ax], fabs(d));
        }
    }
    __device__ __forceinline__ void advance_out(const MarchArgs &g, State &st) const
    {
        st.e += DIM * g.fs;
    }
    // every thread of the CTA calls this (non-emitting lanes carry zeros)
    __device__ __forceinline__ void finish(const MarchArgs &, State &st, int, int, bool) const
    {
        if (!vmax) return;
#pragma unroll
        for (int ax = 0; ax < DIM; ++ax) {
            const double m = warp_max(st.vm[ax]);
            if ((threadIdx.x & 31) == 0 && m > 0.0) atomic_max_nonneg(vmax + ax, m);
        }
    }
};

// ---------------------------------------------------------------------------
// The marching skeleton (register-prefetch version: the fallback of the TMA-fed
// marcher of tma_march.cuh, e.g. for odd grid extents)
// ---------------------------------------------------------------------------
// UNR: unroll the plane loop five times (queue rotates by register renaming);
// otherwise the queue is shifted with moves (smaller code, fewer registers).
// The next plane is prefetched into registers while the current one is staged.
// RINGS = 2: the centre plane kk-2 is shared right before its stencil is formed (store,
//   barrier, neighbour loads: the barrier sits in the middle of the iteration).
// RINGS = 4: every plane is shared as soon as it is staged, into slot kk & 3, and the
//   stencil of plane kk-2 reads the slot written two iterations (two barriers) earlier:
//   the iteration body has no barrier between its stores and its loads, the one barrier
//   sits at its end, and the neighbour loads of plane kk-2 can be issued before / under
//   the staging arithmetic of plane kk (software pipelining across the barrier).
template <int DIM, int TX, int TY, class Op, bool UNR, int RINGS = 2>
struct Marcher {
    using T = TileT<DIM, TX, TY>;
    static constexpr int NF = Op::NF, NPRE = Op::NPRE, NAUX = Op::NAUX;
    static constexpr int RING = RINGS * NF * T::SP;     // doubles; the tables follow
    static_assert(RINGS == 2 || RINGS == 4, "2 or 4 ring slots");
    const MarchArgs &g;
    const DevPhys &P;
    const Op &op;
    double q[NF][5];
    double pre[NPRE];
    double aux[NAUX];
    typename Op::State st;
    int poff, spos, k0, k1, e_aux;
    bool active, emits;

    __device__ __forceinline__ Marcher(const MarchArgs &g_, const DevPhys &P_, const Op &op_)
        : g(g_), P(P_), op(op_)
    {
        const int tid = threadIdx.x;
        if (Op::TABS) {
            // stage the log / exp tables behind the ring
            for (int i = tid; i < KSFD_TAB_DOUBLES; i += T::NT)
                ksfd_smem[RING + i] = __longlong_as_double(
                    (long long)(i < 256 ? g_log_tab[i] : g_exp_tab[i - 256]));
            __syncthreads();
        }
        const int i0 = blockIdx.x * g.ox;
        int x, y = 0;               // tile-relative position incl. halo
        bool interior;
        if (DIM == 2) {
            x = tid;
            active = x < g.ox + 2 * KSFD_SW;
            interior = x >= KSFD_SW && x < g.ox + KSFD_SW;
            spos = x;
            poff = wrapi(i0 - KSFD_SW + x, g.n0);
            emits = interior && (i0 + x - KSFD_SW) < g.n0;
        } else {
            const int j0 = blockIdx.y * g.oy;
            if (tid < TX * TY) {
                const int a = tid % TX, b = tid / TX;
                x = a + KSFD_SW;
                y = b + KSFD_SW;
                active = a < g.ox && b < g.oy;
                interior = active;
            } else {
                int h = tid - TX * TY;
                interior = false;
                if (h < 4 * TX) {               // two rows below, two above
                    const int r = h / TX, a = h - r * TX;
                    x = a + KSFD_SW;
                    y = r < 2 ? r : g.oy + r;
                    active = a < g.ox;
                } else {                        // two columns left, two right
                    h -= 4 * TX;
                    const int cc = h & 3, b = h >> 2;
                    x = cc < 2 ? cc : g.ox + cc;
                    y = b + KSFD_SW;
                    active = b < g.oy && b < TY;    // threads padding the last warp
                }
            }
            spos = y * T::PX + x;
            poff = wrapi(j0 - KSFD_SW + y, g.n1) * g.n0 + wrapi(i0 - KSFD_SW + x, g.n0);
            emits = interior && (i0 + x - KSFD_SW) < g.n0 && (j0 + y - KSFD_SW) < g.n1;
        }
        march_chunk(g, blockIdx.z, k0, k1);
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int s = 0; s < 5; ++s) q[f][s] = 0.0;
    }

    template <int PH>
    __device__ __forceinline__ void step(int kk)
    {
        double cur[NPRE];
#pragma unroll
        for (int c = 0; c < NPRE; ++c) cur[c] = pre[c];
        if (active && kk + 1 < k1 + KSFD_SW) op.load(g, st, kk + 1, poff, RegSink{pre});
        if (active) {
            double f[NF];
            op.stage(P, SmemTabs<RING>(), cur, f);
#pragma unroll
            for (int c = 0; c < NF; ++c) q[c][PH] = f[c];
            if (RINGS == 4) {
                const int wi = (kk & 3) * (NF * T::SP) + spos;
#pragma unroll
                for (int c = 0; c < NF; ++c) ksfd_smem[wi + c * T::SP] = f[c];
            }
        }
        if (RINGS == 4) {
            if (kk - KSFD_SW >= k0) {                   // uniform over the CTA
                if (Op::HAS_AUX && emits) {
                    op.load_aux(g, e_aux, 0, RegSink{aux});
                    e_aux += Op::out_fields(DIM) * g.fs;
                }
                if (emits) {
                    LaneAcc<DIM, NF, T::SP, T::SY, PH> a(q, ((kk - KSFD_SW) & 3) * (NF * T::SP) + spos);
                    op.emit(P, g, a, aux, st);
                }
                op.advance_out(g, st);
            }
            __syncthreads();
            return;
        }
        if (kk - KSFD_SW >= k0) {                       // uniform over the CTA
            if (Op::HAS_AUX && emits) {
                // after the register-hungry stage; the barrier wait hides some latency
                op.load_aux(g, e_aux, 0, RegSink{aux});
                e_aux += Op::out_fields(DIM) * g.fs;
            }
            const int ri = (kk & 1) * (NF * T::SP) + spos;
            if (active) {
#pragma unroll
                for (int c = 0; c < NF; ++c) ksfd_smem[ri + c * T::SP] = q[c][(PH + 3) % 5];
            }
            __syncthreads();
            if (emits) {
                LaneAcc<DIM, NF, T::SP, T::SY, PH> a(q, ri);
                op.emit(P, g, a, aux, st);
            }
            op.advance_out(g, st);
        }
    }

    __device__ __forceinline__ void run()
    {
        int kk = k0 - KSFD_SW;
        const int kend = k1 + KSFD_SW;
        op.init(g, st, kk, k0, poff);
        e_aux = k0 * Op::out_fields(DIM) * g.fs + poff;
        if (active) op.load(g, st, kk, poff, RegSink{pre});
        if (!UNR) {
            for (; kk < kend; ++kk) {
                step<4>(kk);
#pragma unroll
                for (int c = 0; c < NF; ++c) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) q[c][s] = q[c][s + 1];
                }
            }
        } else {
            for (;;) {
                step<0>(kk);
                if (++kk >= kend) break;
                step<1>(kk);
                if (++kk >= kend) break;
                step<2>(kk);
                if (++kk >= kend) break;
                step<3>(kk);
                if (++kk >= kend) break;
                step<4>(kk);
                if (++kk >= kend) break;
            }
        }
        op.finish(g, st, k0, k1, emits);
    }
};

template <class Op, int SP, int RINGS = 2>
constexpr size_t march_smem_bytes()
{
    return sizeof(double) * (RINGS * Op::NF * SP + (Op::TABS ? KSFD_TAB_DOUBLES : 0));
}

template <int DIM, int TX, int TY, class Op, int MINB, bool UNR, int RINGS = 2>
__global__ void __launch_bounds__(TileT<DIM, TX, TY>::NT, MINB)
k_march(const __grid_constant__ MarchArgs g, const __grid_constant__ DevPhys P,
        const __grid_constant__ Op op, const int *__restrict__ skip)
{
    KSFD_PDL_ENTER();
    // pipelined Krylov solver: launched ahead of the convergence test
    if (skip && KSFD_FLAG(skip)) return;
    Marcher<DIM, TX, TY, Op, UNR, RINGS> m(g, P, op);
    m.run();
}
