// Device-side state shared by the linear solvers (blas1_kernels.cuh: pipelined GMRES;
// sweep_op.cuh: the preconditioned Richardson sweeps fused into the marching kernel):
// the all-reduce of a few doubles over NVLink peer memory inside a one-block epilogue,
// the slots of the device state, the host-visible status block.  Device functions and
// structs only (included by several translation units).
#pragma once
#include "device_common.cuh"

// ---------------------------------------------------------------------------
// All-reduce of a few doubles over NVLink peer memory, INSIDE a single-block
// kernel (fused with the reduction / Givens kernels of the solver: no NCCL call
// and no extra launch in the Krylov loop).  Every rank writes its values into
// slot [parity][rank] of every peer's IPC-shared reduce area, publishes the
// exchange number q in the peers' flag words, waits for the flags of all peers
// and sums the contributions in rank order (bitwise identical on all ranks).
// Double-buffered on q&1: exchange q+2 can only start after this rank received
// q+1 from every peer, which they send after they finished reading q.
// ---------------------------------------------------------------------------
#define KSFD_P2P_MAXR 16
#define KSFD_P2P_RED_MAX 72                 // doubles per contribution
#define KSFD_P2P_FLAGS 256                  // flag words at the start of the shared allocation
#define KSFD_P2P_RFLAG0 32                  // reduce flags: word RFLAG0 + source rank
// bounded spin on a flag word written by a peer GPU: gives up after
// KSFD_P2P_TIMEOUT_CYCLES (~2 min at 1.9 GHz: a peer died; ranks that merely skew on
// the host, e.g. while one writes a checkpoint, stay far below it) and raises the
// host-visible error word AND the device-side sticky word `dead` instead of hanging
// the GPU.  Callers do not advance their exchange counter after a failed wait, and
// every later exchange of the context returns at once (`dead`), so no kernel consumes
// a half-received buffer as if it were complete; the host reports the failure at its
// next entry point (ksfd.cu: p2p_check).
#ifndef KSFD_P2P_TIMEOUT_CYCLES
#define KSFD_P2P_TIMEOUT_CYCLES 240000000000ll
#endif
__device__ __forceinline__ bool p2p_spin(volatile unsigned long long *flag,
                                         unsigned long long q, volatile int *err,
                                         volatile unsigned long long *dead = nullptr)
{
    const long long t0 = clock64();
    unsigned spins = 0;
    while (*flag < q) {
        __nanosleep(32);
        if ((++spins & 0xfff) == 0 &&
            (clock64() - t0 > KSFD_P2P_TIMEOUT_CYCLES || (dead && *dead))) {
            if (dead) *dead = 1ull;
            if (err) *err = 1;
            __threadfence_system();
            return false;
        }
    }
    return true;
}

struct P2PRed {
    double *base[KSFD_P2P_MAXR];            // shared allocation of every rank (own included)
    volatile int *err;                      // pinned, host-visible: set when a wait timed out
    long long red_off;                      // doubles from base to the reduce area
    unsigned long long *ctr;                // DEVICE-side exchange counter of this rank
    volatile unsigned long long *dead;      // DEVICE-side sticky word: a peer wait timed out
    int nranks, rank;
};

// vals[0..n) (shared or global memory of this block) <- op over ranks; op 0 = sum, 1 = max
// The exchange number is a DEVICE-side counter, advanced only by exchanges that
// are really made: kernels of the pipelined solver that were launched ahead and
// skip (identically on all ranks: the skip flag is a function of reduced data)
// make no exchange and leave no gap, so the host may launch ahead by different
// amounts on different ranks.
//
// Wire format ("LL", as NCCL's low-latency protocol): every double travels as two 8-byte
// words {32 data bits, 32-bit exchange number}; an aligned 8-byte store is indivisible,
// so a word whose tag equals the current exchange number carries valid data — no
// separate flag, no system-scope fence on either side, one NVLink store latency per
// all-reduce (measured against the fence + flag version: multi-dot tail 9 us -> see
// profiles/r02_multi_gpu_step_breakdown.txt).  Double-buffered on q&1 as before.
__device__ __forceinline__ void p2p_allreduce(const P2PRed &pr, double *vals, int n, int op)
{
    __shared__ unsigned long long q_;
    __shared__ int ok_;
    if (threadIdx.x == 0) {
        q_ = *pr.ctr + 1;
        ok_ = !(pr.dead && *pr.dead);
    }
    __syncthreads();                         // also: vals complete
    if (!ok_) return;                        // uniform: a peer is gone, the host will report it
    const unsigned long long q = q_;
    const unsigned tag = (unsigned)q;
    const int par = (int)(q & 1);
    // words of rank s's contribution in rank r's area: [par][s][2 * KSFD_P2P_RED_MAX]
    const long long slot0 = pr.red_off + (long long)par * pr.nranks * (2 * KSFD_P2P_RED_MAX);
    const int nw = 2 * n;
    for (int t = threadIdx.x; t < nw * pr.nranks; t += blockDim.x) {
        const int r = t / nw, wd = t - r * nw;
        if (r == pr.rank) continue;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[wd >> 1]);
        const unsigned half = (wd & 1) ? (unsigned)(bits >> 32) : (unsigned)bits;
        volatile unsigned long long *dst = reinterpret_cast<volatile unsigned long long *>(
            pr.base[r] + slot0 + (long long)pr.rank * (2 * KSFD_P2P_RED_MAX) + wd);
        *dst = ((unsigned long long)tag << 32) | half;
    }
    __syncthreads();                         // every word is on its way before vals is overwritten
    // collect: thread i < n assembles element i of every rank and reduces in rank order
    // (bitwise identical on all ranks)
    const volatile unsigned long long *area =
        reinterpret_cast<const volatile unsigned long long *>(pr.base[pr.rank] + slot0);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < pr.nranks; ++r) {
            double v;
            if (r == pr.rank) {
                v = vals[i];
            } else {
                const volatile unsigned long long *src = area + (long long)r * (2 * KSFD_P2P_RED_MAX) + 2 * i;
                unsigned long long lo, hi;
                const long long t0 = clock64();
                unsigned spins = 0;
                for (;;) {
                    lo = src[0];
                    hi = src[1];
                    if ((unsigned)(lo >> 32) == tag && (unsigned)(hi >> 32) == tag) break;
                    if ((++spins & 0xfff) == 0 &&
                        (clock64() - t0 > KSFD_P2P_TIMEOUT_CYCLES || (pr.dead && *pr.dead))) {
                        if (pr.dead) *pr.dead = 1ull;
                        if (pr.err) *pr.err = 1;
                        ok_ = 0;
                        break;
                    }
                }
                v = __longlong_as_double((long long)((hi << 32) | (lo & 0xffffffffull)));
            }
            s = r == 0 ? v : (op == 1 ? fmax(s, v) : s + v);
        }
        vals[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0 && ok_) *pr.ctr = q;
}


// ---------------------------------------------------------------------------
// Device state of a solve (`gm`, `gmi`) and the pinned status block the host polls
// ---------------------------------------------------------------------------
#define KSFD_GM_MAXM 64                      // max restart length
#define KSFD_GM_LD (KSFD_GM_MAXM + 1)
// double slots of the device state `gm`
#define GM_H 0                               // H[col*LD + row], rotated in place
#define GM_CS (GM_H + KSFD_GM_LD * KSFD_GM_MAXM)
#define GM_SN (GM_CS + KSFD_GM_MAXM)
#define GM_G (GM_SN + KSFD_GM_MAXM)          // rhs of the least-squares problem
#define GM_Y (GM_G + KSFD_GM_LD)
#define GM_HCOL (GM_Y + KSFD_GM_MAXM)        // raw column of the current step (+ <w,w>)
#define GM_INV (GM_HCOL + KSFD_GM_LD + 1)    // 1/h[j+1,j]
#define GM_BETA (GM_INV + 1)
#define GM_TOL (GM_BETA + 1)
#define GM_CTOL (GM_TOL + 1)
#define GM_RNORM0 (GM_CTOL + 1)
#define GM_RNORM (GM_RNORM0 + 1)
#define GM_DOUBLES (GM_RNORM + 1)
// int slots of `gmi`
#define GMI_CYCLE_DONE 0                     // iteration kernels skip
#define GMI_FINAL 1                          // cycle-start kernels skip
#define GMI_NOUPD 2                          // x-update skips (nothing to add / NaN)
#define GMI_K 3                              // columns of the closed cycle
#define GMI_ITS 4                            // total iterations of the solve
#define GMI_REASON 5
#define GMI_INTS 8

// host-visible progress (pinned, mapped): written by the device, polled by the host
struct GmStatus {
    volatile int seq;            // 2*cycle+1 once the cycle has begun
    volatile int iters_done;     // columns finished in the current cycle
    volatile int cycle_done, final_, reason, its_total;
    volatile int k_cols;         // columns of the cycle that closed (valid with cycle_done)
    volatile int pad_;
    volatile double rnorm, rnorm0;
};

struct GmOpts {
    double rtol, atol, dtol;
    int max_it, m, reorth;
    double cycle_factor;         // close a cycle after this reduction (0: never early)
};

__device__ __forceinline__ double block_sum_partials(const double *partial, int nblocks)
{
    // all threads of one block; returns the sum in every thread
    __shared__ double sm_[32];
    __shared__ double tot_;
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partial[b];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sm_[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < (blockDim.x + 31) / 32; ++q) t += sm_[q];
        tot_ = t;
    }
    __syncthreads();
    return tot_;
}

