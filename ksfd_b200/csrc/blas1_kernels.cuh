// Fused BLAS-1 kernels for the device-resident Krylov solver and the ROSW
// stage combinations (replace PETSc VecMDot / VecMAXPY / VecNorm /
// TSErrorWeightedNorm).  All reductions are two-stage and deterministic:
// per-block partial sums (warp shuffle + shared memory), then one small
// kernel that sums the partials in a fixed order.
#pragma once
#include "device_common.cuh"
#include "solver_state.cuh"
#include "halo_push.cuh"

#define KSFD_RED_BLOCKS 592          // 4 CTAs per SM on 148 SMs
#define KSFD_RED_THREADS 256
#define KSFD_MAXV 8                  // vectors fused per launch

struct VecList {
    const double *v[KSFD_MAXV];
};

// 16-byte (double2) streaming is used when the length is even and every
// pointer is 16-byte aligned (more bytes in flight per thread)
__device__ __forceinline__ bool aligned16(const void *p)
{
    return (reinterpret_cast<unsigned long long>(p) & 15ull) == 0;
}
template <int NV>
__device__ __forceinline__ bool all_aligned16(long long n, const VecList &vs, const void *w)
{
    bool ok = (n & 1) == 0 && aligned16(w);
#pragma unroll
    for (int i = 0; i < NV; ++i) ok = ok && aligned16(vs.v[i]);
    return ok;
}
struct CoefList {
    double c[KSFD_MAXV];
};

template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&acc)[NV],
                                                   double *partial, int stride)
{
    __shared__ double sm[KSFD_RED_THREADS / 32][KSFD_MAXV + 1];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = warp_sum(acc[i]);
        if (l == 0) sm[w][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int q = 0; q < KSFD_RED_THREADS / 32; ++q) s += sm[q][threadIdx.x];
        partial[(size_t)threadIdx.x * stride + blockIdx.x] = s;
    }
}

// partial[i][block] = sum over this block's elements of vs[i]*w   (i < NV)
template <int NV>
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_mdot(long long n, VecList vs, const double *__restrict__ w,
       double *__restrict__ partial)
{
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        const double wv = w[e];
#pragma unroll
        for (int i = 0; i < NV; ++i) acc[i] = fma(__ldg(vs.v[i] + e), wv, acc[i]);
    }
    block_reduce_store<NV>(acc, partial, KSFD_RED_BLOCKS);
}

// out[i] (op)= sum_b partial[i][b];  mode 0: store, 1: add, 2: store sqrt
// (partial and out may alias: in-place sqrt of an all-reduced sum)
__global__ void k_reduce_partials(int nv, int nblocks, const double *partial, double *out,
                                  int mode)
{
    const int i = blockIdx.x;
    if (i >= nv) return;
    __shared__ double sm[32];
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x)
        s += partial[(size_t)i * KSFD_RED_BLOCKS + b];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < (blockDim.x + 31) / 32; ++q) t += sm[q];
        if (mode == 1) out[i] += t;
        else if (mode == 2) out[i] = sqrt(t);
        else out[i] = t;
    }
}

// y = ybase_scale*y + sum_i c[i]*x_i ; coefficients by value (host known)
template <int NV>
__global__ void k_maxpy_host(long long n, CoefList c, VecList xs, double yscale,
                             double *__restrict__ y)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        double s = (yscale == 0.0) ? 0.0 : yscale * y[e];
#pragma unroll
        for (int i = 0; i < NV; ++i) s = fma(c.c[i], __ldg(xs.v[i] + e), s);
        y[e] = s;
    }
}

// w -= sum_i h[i]*V_i  (h on the device), and partial sums of w_new^2
template <int NV>
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_orth_update(long long n, VecList vs, const double *__restrict__ h,
              double *__restrict__ w, double *__restrict__ partial, int want_norm)
{
    double hh[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) hh[i] = h[i];
    double acc[1] = {0.0};
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        double s = w[e];
#pragma unroll
        for (int i = 0; i < NV; ++i) s = fma(-hh[i], __ldg(vs.v[i] + e), s);
        w[e] = s;
        acc[0] = fma(s, s, acc[0]);
    }
    if (want_norm) block_reduce_store<1>(acc, partial, KSFD_RED_BLOCKS);
}

// Gram-Schmidt bookkeeping of one Arnoldi step, one block:
//   h[i] = sum_b partial[i][b]  for i < nv (the last one is <w,w>);
//   then with ALL h[0..j] (earlier batches already reduced into h):
//   hn2 = <w,w> - sum_i h[i]^2 ;  h[j+1] = sqrt(hn2) ;  aux[0] = 1/h[j+1]
//   aux[1] = 1 if the subtraction cancelled too much (caller recomputes the
//   norm explicitly), else 0.
// nv_batch partial rows are reduced into h[off .. off+nv_batch); j+1 = index of
// the <w,w> entry.
__global__ void k_gs_finalize(int nv_batch, int off, int jp1, int nblocks,
                              const double *__restrict__ partial,
                              double *__restrict__ h, double *__restrict__ aux,
                              double thresh)
{
    __shared__ double sums[KSFD_MAXV + 1];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int i = w; i < nv_batch; i += blockDim.x >> 5) {
        double s = 0.0;
        for (int b = l; b < nblocks; b += 32) s += partial[(size_t)i * KSFD_RED_BLOCKS + b];
        s = warp_sum(s);
        if (l == 0) sums[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 0; i < nv_batch; ++i) h[off + i] = sums[i];
        const double ww = h[jp1];
        double ss = 0.0;
        for (int i = 0; i < jp1; ++i) ss = fma(h[i], h[i], ss);
        const double hn2 = ww - ss;
        const bool bad = !(hn2 > thresh * ww);
        const double hn = bad ? 1.0 : sqrt(hn2);
        h[jp1] = bad ? 0.0 : hn;
        aux[0] = 1.0 / hn;
        aux[1] = bad ? 1.0 : 0.0;
    }
}

// same finalisation when the sums are already in h (multi-rank: after the
// all-reduce of the partial dot products)
__global__ void k_gs_finalize_only(int jp1, double *__restrict__ h,
                                   double *__restrict__ aux, double thresh)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const double ww = h[jp1];
        double ss = 0.0;
        for (int i = 0; i < jp1; ++i) ss = fma(h[i], h[i], ss);
        const double hn2 = ww - ss;
        const bool bad = !(hn2 > thresh * ww);
        const double hn = bad ? 1.0 : sqrt(hn2);
        h[jp1] = bad ? 0.0 : hn;
        aux[0] = 1.0 / hn;
        aux[1] = bad ? 1.0 : 0.0;
    }
}

// w = (w - sum_i h[i]*V_i) * scale,  scale = *inv (device) or 1 if inv == NULL
template <int NV>
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_orth_scale(long long n, VecList vs, const double *__restrict__ h,
             const double *__restrict__ inv, double *__restrict__ w)
{
    double hh[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) hh[i] = h[i];
    const double sc = inv ? inv[0] : 1.0;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        double s = w[e];
#pragma unroll
        for (int i = 0; i < NV; ++i) s = fma(-hh[i], __ldg(vs.v[i] + e), s);
        w[e] = s * sc;
    }
}

// y = x * (sign / *norm)     (Krylov vector normalisation; norm on device)
__global__ void k_scale_by_inv(long long n, const double *x,
                               const double *__restrict__ norm, double sign,
                               double *y)
{
    const double f = sign / norm[0];
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x)
        y[e] = x[e] * f;
}

// ROSW stage set-up (PETSc TSStep_RosW):
//   Z = u + sum_j a[j] Y_j ;  Zdot = sum_j g[j] Y_j
// PUSH: several ranks — the boundary planes of Z (the input of the stage residual) also go
// to the neighbours' ghost buffers, first and from blocks of their own (cf. k_gm_orth_scale)
template <int NV, bool PUSH>
__global__ void k_stage_combine(long long n, const double *__restrict__ u,
                                VecList Y, CoefList a, CoefList gm,
                                double *__restrict__ Z, double *__restrict__ Zdot, HaloPush hp);

// ROSW completion + embedded error estimate (TSEvaluateStep_RosW +
// TSErrorWeightedNorm NORM_2):
//   unew = u + sum b[j] Y_j ; uemb = u + sum be[j] Y_j
//   partial += ((unew-uemb)/(atol + rtol*max(|unew|,|uemb|)))^2
template <int NV>
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_complete_step(long long n, const double *__restrict__ u, VecList Y, CoefList b,
                CoefList be, double atol, double rtol, double *__restrict__ unew,
                double *__restrict__ partial)
{
    double acc[1] = {0.0};
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        const double u0 = u[e];
        double un = u0, ue = u0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double y = __ldg(Y.v[i] + e);
            un = fma(b.c[i], y, un);
            ue = fma(be.c[i], y, ue);
        }
        unew[e] = un;
        const double tol = atol + rtol * fmax(fabs(un), fabs(ue));
        const double r = (un - ue) / tol;
        acc[0] = fma(r, r, acc[0]);
    }
    block_reduce_store<1>(acc, partial, KSFD_RED_BLOCKS);
}

// partial sums of field 0 (worm count, KSFD/ksfdts.py:239-246); plane-SoA
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_sum_dof0(long long npts, long long plane_pts, int dof, const double *__restrict__ u,
           double *__restrict__ partial)
{
    double acc[1] = {0.0};
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npts;
         p += (long long)gridDim.x * blockDim.x) {
        const long long k = p / plane_pts;
        acc[0] += u[k * dof * plane_pts + (p - k * plane_pts)];
    }
    block_reduce_store<1>(acc, partial, KSFD_RED_BLOCKS);
}

__global__ void k_scale_dof0(long long npts, long long plane_pts, int dof, double f,
                             double *__restrict__ u)
{
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npts;
         p += (long long)gridDim.x * blockDim.x) {
        const long long k = p / plane_pts;
        u[k * dof * plane_pts + (p - k * plane_pts)] *= f;
    }
}

// rho[p] *= exp(sd * z[p]): lognormal multiplicative noise on the cell density
// (KSFD/ksfdts.py:268-284; z = the standard-normal sample of the rank's numpy stream, one
// value per owned point in x-fastest order, drawn on the host so that the stream is the
// reference's; the field itself never leaves the device)
__global__ void k_mul_exp_dof0(long long npts, long long plane_pts, int dof, double sd,
                               const double *__restrict__ z, double *__restrict__ u)
{
    for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < npts;
         p += (long long)gridDim.x * blockDim.x) {
        const long long k = p / plane_pts;
        u[k * dof * plane_pts + (p - k * plane_pts)] *= exp(sd * z[p]);
    }
}

// stand-alone version for the remaining small reductions (norms, error norm,
// CFL maxima): buf[0..n) in global memory, one block
__global__ void k_p2p_allreduce(P2PRed pr, double *buf, int n, int op)
{
    __shared__ double v[KSFD_P2P_RED_MAX];
    for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] = buf[i];
    p2p_allreduce(pr, v, n, op);
    for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = v[i];
}

// Work split of a pushing producer.  When the boundary positions [0, nb) of the push order
// need only a few blocks (2-D: 24 of 592), those blocks do NOTHING else: they store the
// planes, publish (the publishing thread sits in a system-scope fence for about two NVLink
// round trips) and exit, while the other blocks share the interior — the wait is off the
// kernel's critical path (measured: with statically shared interior work the late block
// extended the kernel by 5-8 us).  Returns false when the boundary needs most of the grid
// (3-D slabs): then every block does boundary and interior work as usual.
struct PushSplit {
    bool split;
    unsigned nbA;           // blocks that own boundary positions
    long long p0, stride;   // first interior position / stride of this thread when split
};
__device__ __forceinline__ unsigned push_blocks(long long nb);
__device__ __forceinline__ PushSplit push_split(bool push, long long nb)
{
    PushSplit ps;
    ps.nbA = push ? push_blocks(nb) : 0;
    ps.split = push && ps.nbA * 4u <= gridDim.x;
    const long long rest = (long long)(gridDim.x - ps.nbA) * blockDim.x;
    ps.stride = ps.split ? rest : (long long)gridDim.x * blockDim.x;
    ps.p0 = nb + ((long long)blockIdx.x - ps.nbA) * blockDim.x + threadIdx.x;
    return ps;
}
// number of blocks that own boundary positions [0, nb) of the push order, and whether this
// block is one of them
__device__ __forceinline__ unsigned push_blocks(long long nb)
{
    const long long need = (nb + blockDim.x - 1) / blockDim.x;
    return (unsigned)(need < (long long)gridDim.x ? need : (long long)gridDim.x);
}

// ---------------------------------------------------------------------------
// Pipelined GMRES: the Hessenberg/Givens bookkeeping, the convergence test and
// the back substitution run on the device; the host only polls a pinned status
// word (no stream synchronisation inside a solve) and may launch ahead.
// Every kernel of the pipeline starts with `if (*skip) return`.
// ---------------------------------------------------------------------------
// Start of a cycle: beta = ||r|| from the partial sums of <r,r> (nblocks of
// them; 1 when already reduced over ranks), tolerances, convergence of the
// TRUE residual.
struct GmBegin {
    int nblocks;                 // partial sums of <r,r>
    const double *partial;
    double *gm;
    int *gmi;
    GmStatus *hs;
    int cycle;
    GmOpts o;
    P2PRed pr;
    unsigned *done;              // block counter of the fused variants
};

// all threads of ONE block
__device__ __forceinline__ void gm_cycle_begin_body(const GmBegin &a)
{
    const double *partial = a.partial;
    double *gm = a.gm;
    int *gmi = a.gmi;
    GmStatus *hs = a.hs;
    const int nblocks = a.nblocks, cycle = a.cycle;
    const GmOpts &o = a.o;
    const P2PRed &pr = a.pr;
    double s = block_sum_partials(partial, nblocks);
    if (pr.nranks > 1) {
        __shared__ double sv[1];
        if (threadIdx.x == 0) sv[0] = s;
        p2p_allreduce(pr, sv, 1, 0);
        s = sv[0];
    }
    if (threadIdx.x != 0) return;
    const double beta = sqrt(s);
    if (cycle == 0) {
        gm[GM_RNORM0] = beta;
        gm[GM_TOL] = fmax(o.rtol * beta, o.atol);
        gmi[GMI_ITS] = 0;
    }
    const double tol = gm[GM_TOL];
    gm[GM_BETA] = beta;
    gm[GM_RNORM] = beta;
    gm[GM_CTOL] = fmax(tol, o.cycle_factor * beta);
    for (int i = 0; i <= o.m; ++i) gm[GM_G + i] = 0.0;
    gm[GM_G] = beta;
    int fin = 0, reason = 0;
    if (!(beta == beta)) { fin = 1; reason = -9; }
    else if (beta == 0.0) { fin = 1; reason = 3; }
    else if (beta <= tol) { fin = 1; reason = 2; }
    else if (gmi[GMI_ITS] >= o.max_it) { fin = 1; reason = -3; }
    gmi[GMI_CYCLE_DONE] = fin;
    gmi[GMI_FINAL] = fin;
    gmi[GMI_NOUPD] = 1;          // until a column exists
    gmi[GMI_K] = 0;
    gmi[GMI_REASON] = reason;
    hs->iters_done = 0;
    hs->k_cols = 0;
    hs->cycle_done = fin;
    hs->final_ = fin;
    hs->reason = reason;
    hs->its_total = gmi[GMI_ITS];
    hs->rnorm = beta;
    hs->rnorm0 = gm[GM_RNORM0];
    __threadfence_system();
    hs->seq = 2 * cycle + 1;
    __threadfence_system();
}

__global__ void k_gm_cycle_begin(GmBegin a)
{
    if (a.gmi[GMI_FINAL]) return;
    gm_cycle_begin_body(a);
}

// the block that finishes last runs gm_cycle_begin_body (same pattern as the
// fused bookkeeping of k_gm_mdot); every thread of the block calls this
__device__ __forceinline__ void gm_begin_in_last_block(const GmBegin &a)
{
    __shared__ int last_;
    __threadfence();                           // this block's partial sum is visible
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(a.done, 1u);
        last_ = (t == gridDim.x - 1);
        if (last_) atomicExch(a.done, 0u);
    }
    __syncthreads();
    if (!last_) return;
    __threadfence();                           // see every block's partial sum
    gm_cycle_begin_body(a);
}

// <r,r> of the right-hand side and the start of cycle 0 in one launch
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_gm_norm_begin(long long n, const double *__restrict__ r, double *__restrict__ partial,
                GmBegin a)
{
    KSFD_PDL_ENTER();
    double acc[1] = {0.0};
    if ((n & 1) == 0 && aligned16(r)) {
        const double2 *r2 = reinterpret_cast<const double2 *>(r);
        for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (n >> 1);
             e += (long long)gridDim.x * blockDim.x) {
            const double2 v = r2[e];
            acc[0] = fma(v.y, v.y, fma(v.x, v.x, acc[0]));
        }
    } else {
        for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
             e += (long long)gridDim.x * blockDim.x) {
            const double v = r[e];
            acc[0] = fma(v, v, acc[0]);
        }
    }
    block_reduce_store<1>(acc, partial, KSFD_RED_BLOCKS);
    gm_begin_in_last_block(a);
}

// Element order of the producers that push: the two boundary plane pairs FIRST (the
// first 2*cnt elements of the order), then the interior.  The peer stores go out at the
// start of the kernel and are acknowledged while the same threads stream the interior,
// so the system-scope fence before the flag is published (halo_push_publish) finds
// nothing outstanding instead of adding an NVLink round trip to the kernel's tail.  Maps position p of the order to the element index (units: elements,
// or element pairs when everything is counted in double2).
__device__ __forceinline__ long long push_order(long long p, long long cnt, long long n)
{
    if (p < cnt) return p;                          // bottom planes
    if (p < 2 * cnt) return n - 2 * cnt + p;        // top planes
    return p - cnt;                                 // interior: cnt .. n - cnt
}

// y = x * (sign / gm[GM_BETA])    (first Krylov vector)
// PUSH: the halo push of y is fused in (several ranks over peer memory)
template <bool PUSH>
__global__ void k_gm_first_vector(long long n, const double *x, const double *__restrict__ gm,
                                  const int *__restrict__ gmi, double sign, double *y,
                                  HaloPush hp)
{
    KSFD_PDL_ENTER();
    if (KSFD_FLAG(gmi + GMI_FINAL)) return;
    const double f = sign / KSFD_FLAG(gm + GM_BETA);
    const bool push = PUSH && halo_push_on(hp);
    unsigned long long q = 0;
    const long long sh = push ? halo_push_shift(hp, q) : 0;
    // (with nloc < 4 the plane pairs overlap: every position is a boundary position)
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if ((n & 1) == 0 && aligned16(x) && aligned16(y)) {
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        double2 *y2 = reinterpret_cast<double2 *>(y);
        const long long n2 = n >> 1, c2 = hp.cnt >> 1;
        const bool reorder = push && 2 * c2 <= n2;
        const long long nb = push ? (reorder ? 2 * c2 : n2) : 0;     // boundary positions
        for (; p < nb; p += stride) {
            const long long e = reorder ? push_order(p, c2, n2) : p;
            double2 v = x2[e];
            v.x *= f;
            v.y *= f;
            y2[e] = v;
            halo_push2(hp, sh, e, v);
        }
        const PushSplit ps = push_split(push, nb);
        if (push && blockIdx.x < ps.nbA) halo_push_publish(hp, q, ps.nbA);
        if (ps.split) {
            if (blockIdx.x < ps.nbA) return;
            p = ps.p0;
        }
        for (; p < n2; p += ps.stride) {
            const long long e = reorder ? push_order(p, c2, n2) : p;
            double2 v = x2[e];
            v.x *= f;
            v.y *= f;
            y2[e] = v;
        }
    } else {
        const bool reorder = push && 2 * hp.cnt <= n;
        const long long nb = push ? (reorder ? 2 * hp.cnt : n) : 0;
        for (; p < nb; p += stride) {
            const long long e = reorder ? push_order(p, hp.cnt, n) : p;
            const double v = x[e] * f;
            y[e] = v;
            halo_push1(hp, sh, e, v);
        }
        const PushSplit ps = push_split(push, nb);
        if (push && blockIdx.x < ps.nbA) halo_push_publish(hp, q, ps.nbA);
        if (ps.split) {
            if (blockIdx.x < ps.nbA) return;
            p = ps.p0;
        }
        for (; p < n; p += ps.stride) {
            const long long e = reorder ? push_order(p, hp.cnt, n) : p;
            y[e] = x[e] * f;
        }
    }
}

// partial[i][block] = <vs[i], w>, skipping when the cycle is closed
// hcol[off + i] = sum_b partial[i][b], i < nv  (earlier batches of a long column)
__global__ void k_gm_reduce(int nv, int off, int nblocks, const double *__restrict__ partial,
                            const int *__restrict__ gmi, double *__restrict__ gm)
{
    if (gmi[GMI_CYCLE_DONE]) return;
    const int i = blockIdx.x;
    if (i >= nv) return;
    const double s = block_sum_partials(partial + (size_t)i * KSFD_RED_BLOCKS, nblocks);
    if (threadIdx.x == 0) gm[GM_HCOL + off + i] = s;
}

// End of Arnoldi step j (one block).  Reduces the last batch of partial sums
// (nv_batch rows into hcol[off..]; nv_batch = 0 when hcol is complete already,
// e.g. after the all-reduce over ranks), forms h[j+1] from <w,w> - sum h_i^2,
// applies the Givens rotations, updates the residual norm, decides whether the
// cycle / the solve ends, and if the cycle ends solves the triangular system.
// The serial part works out of shared memory (every global access of a single
// thread costs an L2 round trip).
struct GmFin {                      // arguments of the end-of-step bookkeeping
    int nv_batch, off, j, nblocks;
    const double *partial;
    double *gm;
    int *gmi;
    GmStatus *hs;
    GmOpts o;
    P2PRed pr;
    unsigned *done;                 // block counter when fused into the multi-dot kernel
};

// executed by ONE block of 256 threads
__device__ __forceinline__ void gm_finalize_body(const GmFin &a)
{
    const int nv_batch = a.nv_batch, off = a.off, j = a.j, nblocks = a.nblocks;
    const double *__restrict__ partial = a.partial;
    double *__restrict__ gm = a.gm;
    int *__restrict__ gmi = a.gmi;
    GmStatus *hs = a.hs;
    const GmOpts &o = a.o;
    const P2PRed &pr = a.pr;
    __shared__ double h[KSFD_GM_LD + 1], cs[KSFD_GM_MAXM], sn[KSFD_GM_MAXM], y[KSFD_GM_MAXM];
    __shared__ double sc[6];            // g[j], tol, ctol, rnorm0, its(as double), -
    __shared__ double Ht[KSFD_GM_MAXM * (KSFD_GM_MAXM + 1) / 2];
    __shared__ double gs[KSFD_GM_LD];
    __shared__ int dec[4];              // cyc, fin, reason, noupd
    const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    // columns reduced by earlier batches (or all of them when nv_batch == 0)
    for (int i = tid; i <= j + 1; i += blockDim.x)
        if (i < off || nv_batch == 0) h[i] = gm[GM_HCOL + i];
    for (int i = tid; i < j; i += blockDim.x) {
        cs[i] = gm[GM_CS + i];
        sn[i] = gm[GM_SN + i];
    }
    if (tid == 0) {
        sc[0] = gm[GM_G + j];
        sc[1] = gm[GM_TOL];
        sc[2] = gm[GM_CTOL];
        sc[3] = gm[GM_RNORM0];
        sc[4] = (double)gmi[GMI_ITS];
    }
    for (int i = w; i < nv_batch; i += blockDim.x >> 5) {
        double s = 0.0;
        for (int b = l; b < nblocks; b += 32) s += partial[(size_t)i * KSFD_RED_BLOCKS + b];
        s = warp_sum(s);
        if (l == 0) h[off + i] = s;
    }
    __syncthreads();
    if (pr.nranks > 1) p2p_allreduce(pr, h, j + 2, 0);      // sum the column over the ranks
    const int k = j + 1;
    if (tid == 0) {
        const double ww = h[j + 1];
        double ss = 0.0;
        for (int i = 0; i <= j; ++i) ss = fma(h[i], h[i], ss);
        const double hn2 = ww - ss;
        // <w,w> - sum h^2 carries an absolute error ~eps*<w,w>: below 1e-8*<w,w>
        // the norm of the new direction is no longer reliable (it lies,
        // numerically, inside the current space).  Keep the column with the
        // best available norm, do not normalise w, and close the cycle: the
        // next one restarts from the TRUE residual, which also decides
        // convergence.
        const bool bad = !(hn2 > 1e-8 * ww);
        const double hn = hn2 > 0.0 ? sqrt(hn2) : 0.0;
        gm[GM_INV] = (bad || hn == 0.0) ? 0.0 : 1.0 / hn;
        // keep the raw (rank-summed) column for the orthogonalisation kernel
        if (pr.nranks > 1) {
            for (int i = 0; i <= j + 1; ++i) gm[GM_HCOL + i] = h[i];
        } else {
            for (int i = 0; i < nv_batch; ++i) gm[GM_HCOL + off + i] = h[off + i];
        }
        h[j + 1] = hn;
        for (int i = 0; i < j; ++i) {
            const double t = cs[i] * h[i] + sn[i] * h[i + 1];
            h[i + 1] = -sn[i] * h[i] + cs[i] * h[i + 1];
            h[i] = t;
        }
        const double den = hypot(h[j], h[j + 1]);
        const double c_ = den > 0.0 ? h[j] / den : 1.0, s_ = den > 0.0 ? h[j + 1] / den : 0.0;
        h[j] = den;
        h[j + 1] = 0.0;
        gm[GM_CS + j] = c_;
        gm[GM_SN + j] = s_;
        const double gj = sc[0];
        gm[GM_G + j + 1] = -s_ * gj;
        gm[GM_G + j] = c_ * gj;
        const double rnorm = fabs(s_ * gj);
        gm[GM_RNORM] = rnorm;
        const int its = (int)sc[4] + 1;
        gmi[GMI_ITS] = its;
        const double tol = sc[1], ctol = sc[2];
        int fin = 0, cyc = 0, reason = 0, noupd = 0;
        if (!(rnorm == rnorm)) { fin = cyc = 1; reason = -9; noupd = 1; }
        else if (bad) cyc = 1;
        else if (rnorm <= tol && ctol <= tol) { fin = cyc = 1; reason = 2; }
        else if (o.dtol > 0.0 && rnorm > o.dtol * sc[3]) { fin = cyc = 1; reason = -4; }
        else if (its >= o.max_it) { fin = cyc = 1; reason = -3; }
        else if (rnorm <= ctol || j == o.m - 1) cyc = 1;
        dec[0] = cyc;
        dec[1] = fin;
        dec[2] = reason;
        dec[3] = noupd;
        sc[5] = rnorm;
        if (!cyc) hs->iters_done = k;       // the common case: one word, no fence
    }
    __syncthreads();
    // rotated column -> H (global), by all threads
    for (int i = tid; i <= j + 1; i += blockDim.x) gm[GM_H + (size_t)j * KSFD_GM_LD + i] = h[i];
    if (!dec[0]) return;
    // the cycle ends: back substitution on the k x k triangle (packed in smem)
    if (!dec[3]) {
        __syncthreads();
        for (int e = tid; e < k * (k + 1) / 2; e += blockDim.x) {
            // e -> (col q, row i <= q), packed by columns
            int q = (int)((sqrt(8.0 * e + 1.0) - 1.0) * 0.5);
            while (q * (q + 1) / 2 > e) --q;
            while ((q + 1) * (q + 2) / 2 <= e) ++q;
            const int i = e - q * (q + 1) / 2;
            Ht[e] = (q == j) ? h[i] : gm[GM_H + (size_t)q * KSFD_GM_LD + i];
        }
        for (int i = tid; i <= k; i += blockDim.x) gs[i] = gm[GM_G + i];
        __syncthreads();
        if (tid == 0) {
            // g[j], g[j+1] were written above by this thread: use the values
            gs[j] = gm[GM_G + j];
            for (int i = k - 1; i >= 0; --i) {
                double s = gs[i];
                for (int q = i + 1; q < k; ++q) s -= Ht[q * (q + 1) / 2 + i] * y[q];
                y[i] = s / Ht[i * (i + 1) / 2 + i];
            }
        }
        __syncthreads();
        for (int i = tid; i < k; i += blockDim.x) gm[GM_Y + i] = y[i];
    }
    __syncthreads();
    if (tid == 0) {
        gmi[GMI_K] = k;
        gmi[GMI_NOUPD] = dec[3];
        gmi[GMI_REASON] = dec[2];
        gmi[GMI_FINAL] = dec[1];
        __threadfence();
        gmi[GMI_CYCLE_DONE] = 1;
        // order matters: a host that sees iters_done == k must also see
        // cycle_done (with several ranks a step launched on one rank only
        // would desynchronise their NCCL call sequences)
        hs->rnorm = sc[5];
        hs->its_total = gmi[GMI_ITS];
        hs->reason = dec[2];
        hs->final_ = dec[1];
        hs->k_cols = k;
        __threadfence_system();
        hs->cycle_done = 1;
        __threadfence_system();
        hs->iters_done = k;
    }
}

__global__ void k_gm_finalize(GmFin a)
{
    if (a.gmi[GMI_CYCLE_DONE]) return;
    gm_finalize_body(a);
}

// FUSE: the block that finishes last also does the end-of-step bookkeeping
// (gm_finalize_body): no separate one-block kernel per Arnoldi step.
template <int NV, bool FUSE>
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_gm_mdot(long long n, VecList vs, const double *__restrict__ w,
          const int *__restrict__ gmi, double *__restrict__ partial, GmFin fin)
{
    KSFD_PDL_ENTER();
    if (KSFD_FLAG(gmi + GMI_CYCLE_DONE)) return;
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.0;
    if (all_aligned16<NV>(n, vs, w)) {
        const double2 *w2 = reinterpret_cast<const double2 *>(w);
        for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (n >> 1);
             e += (long long)gridDim.x * blockDim.x) {
            const double2 wv = w2[e];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const double2 a = __ldg(reinterpret_cast<const double2 *>(vs.v[i]) + e);
                acc[i] = fma(a.y, wv.y, fma(a.x, wv.x, acc[i]));
            }
        }
    } else {
        for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
             e += (long long)gridDim.x * blockDim.x) {
            const double wv = w[e];
#pragma unroll
            for (int i = 0; i < NV; ++i) acc[i] = fma(__ldg(vs.v[i] + e), wv, acc[i]);
        }
    }
    block_reduce_store<NV>(acc, partial, KSFD_RED_BLOCKS);
    if (FUSE) {
        __shared__ int last_;
        __threadfence();                       // this block's partial sums are visible
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned t = atomicAdd(fin.done, 1u);
            last_ = (t == gridDim.x - 1);
            if (last_) atomicExch(fin.done, 0u);
        }
        __syncthreads();
        if (!last_) return;
        __threadfence();                       // see every block's partial sums
        gm_finalize_body(fin);
    }
}


// w = (w - sum_i h[i]*V_i) * inv   with h, inv from the device state
// PUSH: fused halo push of the finished vector (last batch of a long column only)
template <int NV, bool PUSH>
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_gm_orth_scale(long long n, VecList vs, int off, int do_scale,
                const double *__restrict__ gm, const int *__restrict__ gmi,
                double *__restrict__ w, HaloPush hp)
{
    // the step that closed the cycle does not need its new basis vector
    KSFD_PDL_ENTER();
    if (KSFD_FLAG(gmi + GMI_CYCLE_DONE)) return;
    double hh[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) hh[i] = KSFD_FLAG(gm + GM_HCOL + off + i);
    const double sc = do_scale ? KSFD_FLAG(gm + GM_INV) : 1.0;
    const bool push = PUSH && halo_push_on(hp);
    unsigned long long q = 0;
    const long long sh = push ? halo_push_shift(hp, q) : 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (all_aligned16<NV>(n, vs, w)) {
        double2 *w2 = reinterpret_cast<double2 *>(w);
        const long long n2 = n >> 1, c2 = hp.cnt >> 1;
        const bool reorder = push && 2 * c2 <= n2;      // boundary planes first (push_order)
        const long long nb = push ? (reorder ? 2 * c2 : n2) : 0;
        auto body = [&](long long e) {
            double2 s = w2[e];
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const double2 a = __ldg(reinterpret_cast<const double2 *>(vs.v[i]) + e);
                s.x = fma(-hh[i], a.x, s.x);
                s.y = fma(-hh[i], a.y, s.y);
            }
            s.x *= sc;
            s.y *= sc;
            w2[e] = s;
            return s;
        };
        for (; p < nb; p += stride) {
            const long long e = reorder ? push_order(p, c2, n2) : p;
            halo_push2(hp, sh, e, body(e));
        }
        const PushSplit ps = push_split(push, nb);
        if (push && blockIdx.x < ps.nbA) halo_push_publish(hp, q, ps.nbA);
        if (ps.split) {
            if (blockIdx.x < ps.nbA) return;
            p = ps.p0;
        }
        for (; p < n2; p += ps.stride) body(reorder ? push_order(p, c2, n2) : p);
    } else {
        const bool reorder = push && 2 * hp.cnt <= n;
        const long long nb = push ? (reorder ? 2 * hp.cnt : n) : 0;
        auto body = [&](long long e) {
            double s = w[e];
#pragma unroll
            for (int i = 0; i < NV; ++i) s = fma(-hh[i], __ldg(vs.v[i] + e), s);
            s *= sc;
            w[e] = s;
            return s;
        };
        for (; p < nb; p += stride) {
            const long long e = reorder ? push_order(p, hp.cnt, n) : p;
            halo_push1(hp, sh, e, body(e));
        }
        const PushSplit ps = push_split(push, nb);
        if (push && blockIdx.x < ps.nbA) halo_push_publish(hp, q, ps.nbA);
        if (ps.split) {
            if (blockIdx.x < ps.nbA) return;
            p = ps.p0;
        }
        for (; p < n; p += ps.stride) body(reorder ? push_order(p, hp.cnt, n) : p);
    }
}

// r = sign*rhs - Ax (in place in ax) with the partial sums of <r,r>
// FUSE: the last block also starts the cycle (gm_cycle_begin_body)
template <bool FUSE>
__global__ void __launch_bounds__(KSFD_RED_THREADS)
k_gm_true_residual(long long n, const double *__restrict__ rhs, double sign,
                   const int *__restrict__ gmi, double *__restrict__ ax,
                   double *__restrict__ partial, GmBegin a)
{
    KSFD_PDL_ENTER();
    if (KSFD_FLAG(gmi + GMI_FINAL)) return;
    double acc[1] = {0.0};
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n;
         e += (long long)gridDim.x * blockDim.x) {
        const double r = fma(sign, rhs[e], -ax[e]);
        ax[e] = r;
        acc[0] = fma(r, r, acc[0]);
    }
    block_reduce_store<1>(acc, partial, KSFD_RED_BLOCKS);
    if (FUSE) gm_begin_in_last_block(a);
}

template <int NV, bool PUSH>
__global__ void k_stage_combine(long long n, const double *__restrict__ u,
                                VecList Y, CoefList a, CoefList gm,
                                double *__restrict__ Z, double *__restrict__ Zdot, HaloPush hp)
{
    const bool push = PUSH && halo_push_on(hp);
    unsigned long long q = 0;
    const long long sh = push ? halo_push_shift(hp, q) : 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool reorder = push && 2 * hp.cnt <= n;
    const long long nb = push ? (reorder ? 2 * hp.cnt : n) : 0;     // boundary positions
    auto body = [&](long long e) {
        double z = u[e], zd = 0.0;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const double y = __ldg(Y.v[i] + e);
            z = fma(a.c[i], y, z);
            zd = fma(gm.c[i], y, zd);
        }
        Z[e] = z;
        Zdot[e] = zd;
        return z;
    };
    for (; p < nb; p += stride) {
        const long long e = reorder ? push_order(p, hp.cnt, n) : p;
        halo_push1(hp, sh, e, body(e));
    }
    const PushSplit ps = push_split(push, nb);
    if (push && blockIdx.x < ps.nbA) halo_push_publish(hp, q, ps.nbA);
    if (ps.split) {
        if (blockIdx.x < ps.nbA) return;
        p = ps.p0;
    }
    for (; p < n; p += ps.stride) body(reorder ? push_order(p, hp.cnt, n) : p);
}
