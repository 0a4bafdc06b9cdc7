// TMA-fed marching stencil kernels (sm_100a): the data movement of the marching
// skeleton of march_kernels.cuh re-done around the Blackwell copy engine.
//
// What changes against k_march (the operator policies — stage / emit arithmetic —
// are shared, results are bit-identical):
//   * raw input planes arrive in shared memory by cp.async.bulk.tensor (TMA, SASS
//     UTMALDG) issued by ONE elected thread and completed on mbarriers: no
//     per-thread address arithmetic, no prefetch registers, SC planes in flight
//     per CTA independent of the warps' progress;
//   * there are no halo-lane threads.  A CTA has exactly one thread per OUTPUT
//     point of its TX x TY tile; the 4*TX + 4*TY halo points of the centre plane
//     (the x/y neighbours of the tile edge) are staged by the first warps in an
//     extra pass from their own TMA boxes, which lag the centre boxes by the two
//     planes of the register queue.  Threads that only ever staged (1/3 of a
//     16x16 CTA, each holding a full register queue) are gone: twice as many
//     output points are resident per SM;
//   * the tile is fetched as 5 boxes per input vector and plane — centre TX x TY,
//     two y strips TX x 2, two x strips 2 x TY — each with its start coordinate
//     wrapped periodically, so no box straddles the periodic boundary and the
//     same code serves every tile (corners are never needed: star stencil);
//   * the last tile of an axis is CLAMPED to end at the boundary instead of being
//     partial (it recomputes a few outputs of its neighbour, writing identical
//     values), so every tile is full and lies inside the domain.
//
// Tensor maps (host: tma_host.h; passed in the kernel parameter block): every input
// vector is described as a rank-3 tensor (x, y, plane*nc + field) over its plane-SoA
// storage; coordinate 2 of a box = kofs + k*nc selects the nc fields of plane k at once.
//
// Shared memory (doubles): [staged ring 2 x NF x SP][log/exp tables]
//   [centre ring SC x NPRE x TX*TY][halo ring SH x NPRE x NH][mbarriers SC + SH]
#pragma once
#include <cuda.h>

#include "march_kernels.cuh"

// One input vector: the owned planes live in one buffer (map set 0), the four ghost planes
// of the last axis either are periodic images of owned planes (wrap, one rank) or live in a
// second buffer (map set 1): the halo slot the neighbours fill (several ranks; double-
// buffered on the parity of a device-side exchange counter), or the same buffer (the
// coefficient field, which is stored ghosted).
struct TmaVecIn {
    int kofs[3];    // coordinate 2 of: [0] plane 0 in set 0, [1] plane -2 in set 1, [2] plane nloc in set 1
    int wrap;       // 1: ghost planes = periodic images of the owned planes
    const unsigned long long *par;      // exchange counter (nullptr: none)
    int parshift;   // added to kofs[1], kofs[2] when the counter is odd
    int pad_;
    int nc, coff;   // fields per plane; index of the first one among the operator's raw fields
    // Several ranks over NVLink peer memory: the neighbours store the ghost planes into
    // this rank's halo slot and then publish the exchange number in flag_lo / flag_hi.
    // The producer of the vector only PUSHES (ksfd.cu: k_halo_push, or fused into the
    // BLAS-1 kernel that writes the vector); the wait happens here, in the elected
    // thread of the CTAs that fetch ghost planes, right before their first ghost box —
    // CTAs that own interior chunks never wait, and the NVLink round trip hides behind
    // the owned planes.  nullptr: nothing to wait for.
    const volatile unsigned long long *flag_lo, *flag_hi;
    volatile int *err;                      // host-visible: a wait timed out
    volatile unsigned long long *dead;      // device-side sticky copy of it
};
// Where the tensor maps live: inside the kernel parameter block (__grid_constant__, the
// default: nothing to allocate or copy) or in global memory (-DKSFD_TMAP_PARAM=0, a device
// pool owned by the caller).  Measured on B200 with otherwise identical kernels
// (profiles/r02_tma_tuner_runs.txt, runs 4g / 4p): no difference.
#ifndef KSFD_TMAP_PARAM
#define KSFD_TMAP_PARAM 1
#endif
template <int NIN>
struct TmaInT {
#if KSFD_TMAP_PARAM
    CUtensorMap m[NIN][2][3];   // [input vector][map set][box shape: centre, y strip, x strip]
    __device__ __forceinline__ const CUtensorMap *maps(int i, int set) const { return &m[i][set][0]; }
#else
    const CUtensorMap *m[NIN][2];   // device memory: the three box shapes of [vector][map set]
    __device__ __forceinline__ const CUtensorMap *maps(int i, int set) const { return m[i][set]; }
#endif
    TmaVecIn v[NIN];
};

namespace ktma {
__device__ __forceinline__ unsigned s32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KSFD_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KSFD_DONE;\n"
        "bra KSFD_WAIT;\n"
        "KSFD_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// one box of a rank-3 tensor -> shared memory, completion on an mbarrier
__device__ __forceinline__ void load3(unsigned dst, const CUtensorMap *m, int c0, int c1, int c2,
                                      unsigned bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
}   // namespace ktma

// layout constants shared by the marcher and the issuer
template <int DIM, int TX, int TY, class Op, int SC, int SH>
struct TmaLayout {
    using T = TileT<DIM, TX, TY>;
    static constexpr int NTH = (DIM == 2) ? TX : TX * TY;           // threads = outputs
    static constexpr int NH = (DIM == 2) ? 4 : 4 * TX + 4 * TY;     // halo points per plane
    static constexpr int NF = Op::NF, NPRE = Op::NPRE, NIN = Op::NIN;
    static constexpr int r16(int x) { return (x + 15) / 16 * 16; }
    static constexpr int RING = r16(2 * NF * T::SP);
    static constexpr int TABS = Op::TABS ? KSFD_TAB_DOUBLES : 0;
    static constexpr int CRING = RING + TABS;
    static constexpr int CSLOT = NPRE * NTH;
    static constexpr int HRING = CRING + SC * CSLOT;
    static constexpr int HVEC2 = 16;        // 2-D: doubles reserved per (vector, side) box
    static constexpr int HSLOT = (DIM == 2) ? NIN * 2 * HVEC2 : NPRE * NH;
    static constexpr int HBYTES = (DIM == 2) ? NPRE * 4 * 8 : NPRE * NH * 8;
    static constexpr int BAR = HRING + SH * HSLOT;
    static constexpr int SMEM_DOUBLES = BAR + SC + SH;
    // halo slot: offset of the box of (vector i, piece p)
    //   3-D pieces: 0 bottom rows, 1 top rows, 2 left columns, 3 right columns
    //   2-D pieces: 0 left, 1 right
    __host__ __device__ static constexpr int hbox(int i, int p)
    {
        if (DIM == 2) return (2 * i + p) * HVEC2;
        const int po = p == 0 ? 0 : p == 1 ? 2 * TX : p == 2 ? 4 * TX : 4 * TX + 2 * TY;
        return Op::coff(i) * NH + Op::nc(i) * po;
    }
};

// The neighbour's ghost planes have landed: bounded spin on the flag word it publishes
// after its stores (blas1_kernels.cuh: p2p_spin), then order the TMA reads (async proxy)
// after the observation.  Out of line, arguments by value: only reached with several ranks.
static __device__ __noinline__ void tma_halo_wait(const volatile unsigned long long *f,
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
    asm volatile("fence.proxy.async;" ::: "memory");
}

template <int DIM, int TX, int TY, class Op, bool UNR, int SC, int SH>
struct TmaMarcher {
    using T = TileT<DIM, TX, TY>;
    using L = TmaLayout<DIM, TX, TY, Op, SC, SH>;
    static constexpr int NTH = L::NTH, NH = L::NH;
    static constexpr int NF = Op::NF, NPRE = Op::NPRE, NAUX = Op::NAUX, NIN = Op::NIN;
    static constexpr int RING = L::RING, CRING = L::CRING, CSLOT = L::CSLOT, HRING = L::HRING;
    static constexpr int HSLOT = L::HSLOT, BAR = L::BAR, SMEM_DOUBLES = L::SMEM_DOUBLES;
    static_assert(NTH % 32 == 0 && L::TABS % 16 == 0, "alignment of the TMA destinations");
    static_assert(DIM == 2 || (TX % 8 == 0 && TY % 8 == 0), "3-D tiles: multiples of 8");
    static_assert(NH <= NTH, "halo pass: one halo point per thread");
    __device__ static constexpr int hbox(int i, int p) { return L::hbox(i, p); }

    const MarchArgs &g;
    const DevPhys &P;
    const Op &op;
    const TmaInT<Op::NIN> &tin;
    // Issuers: input vector i is fetched by lane 0 of warp NW-1-i — the LAST warps, which
    // have slack while the first ones stage the halo points — so the ~100 instructions of
    // coordinate set-up and TMA issues per vector and plane run in parallel instead of all
    // on thread 0 (which also runs the halo pass).  Every issuer arrives on the slot's
    // mbarrier with the byte count of its own boxes (arrival count NIN).
    int ivec;                   // input vector this thread fetches, -1: none
    int ipshift;                // parity shift of its ghost planes
    unsigned long long ixq;     // exchange number its ghost planes must have reached
    unsigned iwaited;           // bit 0 / 1: lo / hi flag seen
    double q[NF][5];
    double aux[NAUX];
    typename Op::State st;
    int spos, k0, k1, poff_;
    int hspos, hfs, hoff[NIN];
    int cs, hs;                 // ring slots of the next centre / halo plane to consume
    unsigned cph, hph;          // their mbarrier phase parities
    unsigned bar0;              // shared address of the first mbarrier
    bool active, hact;

    __device__ __forceinline__ int tile_x0() const { return min((int)blockIdx.x * g.ox, g.n0 - g.ox); }
    __device__ __forceinline__ int tile_y0() const
    {
        return DIM == 2 ? 0 : min((int)blockIdx.y * g.oy, g.n1 - g.oy);
    }

    __device__ __forceinline__ TmaMarcher(const MarchArgs &g_, const DevPhys &P_, const Op &op_,
                                          const TmaInT<Op::NIN> &tin_)
        : g(g_), P(P_), op(op_), tin(tin_)
    {
        const int tid = threadIdx.x;
        bar0 = ktma::s32(ksfd_smem + BAR);
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < SC + SH; ++s) ktma::mbar_init(bar0 + 8 * s, NIN);
            ktma::mbar_fence_init();
        }
        if (Op::TABS) {
            for (int i = tid; i < KSFD_TAB_DOUBLES; i += NTH)
                ksfd_smem[RING + i] = __longlong_as_double(
                    (long long)(i < 256 ? g_log_tab[i] : g_exp_tab[i - 256]));
        }
        const int i0 = tile_x0(), j0 = tile_y0();
        int poff;
        hact = false;
        hspos = 0;
        hfs = 2;
#pragma unroll
        for (int i = 0; i < NIN; ++i) hoff[i] = 0;
        if (DIM == 2) {
            active = tid < g.ox;
            spos = tid + KSFD_SW;
            poff = i0 + tid;
            if (tid < 4) {
                const int p = tid >> 1, cc = tid & 1;
                hact = true;
                hspos = p ? g.ox + KSFD_SW + cc : cc;
#pragma unroll
                for (int i = 0; i < NIN; ++i) hoff[i] = (p ? hbox(i, 1) : hbox(i, 0)) + cc;
            }
        } else {
            const int a = tid % TX, b = tid / TX;
            active = a < g.ox && b < g.oy;
            spos = (b + KSFD_SW) * T::PX + a + KSFD_SW;
            poff = (j0 + b) * g.n0 + i0 + a;
            if (tid < NH) {
                int p, idx;
                if (tid < 4 * TX) {                     // rows below (p 0) / above (p 1) the tile
                    p = tid >= 2 * TX;
                    idx = tid - p * 2 * TX;
                    const int r = idx / TX, aa = idx - r * TX;
                    hspos = (p ? g.oy + KSFD_SW + r : r) * T::PX + aa + KSFD_SW;
                    hact = aa < g.ox;
                    hfs = 2 * TX;
                } else {                                // columns left (p 2) / right (p 3)
                    const int h = tid - 4 * TX;
                    p = 2 + (h >= 2 * TY);
                    idx = h - (p - 2) * 2 * TY;
                    const int bb = idx >> 1, cc = idx & 1;
                    hspos = (bb + KSFD_SW) * T::PX + (p == 3 ? g.ox + KSFD_SW + cc : cc);
                    hact = bb < g.oy;
                    hfs = 2 * TY;
                }
#pragma unroll
                for (int i = 0; i < NIN; ++i)
                    hoff[i] = (p == 0 ? hbox(i, 0) : p == 1 ? hbox(i, 1) : p == 2 ? hbox(i, 2)
                                                                                 : hbox(i, 3)) +
                              idx;
            }
        }
        poff_ = poff;
        march_chunk(g, blockIdx.z, k0, k1);
        op.init_out(g, st, k0, poff);
        if constexpr (Op::NEEDS_OWNER) {
            // the clamped last tile of an axis overlaps its neighbour: the outputs in front
            // of its unclamped origin belong to the neighbour
            bool own = active && (DIM == 2 ? tid : tid % TX) >= (int)blockIdx.x * g.ox - i0;
            if (DIM == 3) own = own && tid / TX >= (int)blockIdx.y * g.oy - j0;
            st.own = own;
        }
        cs = hs = 0;
        cph = hph = 0;
#pragma unroll
        for (int f = 0; f < NF; ++f)
#pragma unroll
            for (int s = 0; s < 5; ++s) q[f][s] = 0.0;
        __syncthreads();                    // barriers initialised, tables staged
        {
            constexpr int NW = NTH / 32;
            const int w = tid >> 5;
            static_assert(NW >= NIN, "one issuing warp per input vector");
            ivec = ((tid & 31) == 0 && NW - 1 - w < NIN) ? NW - 1 - w : -1;
            ipshift = 0;
            ixq = 0ull;
            iwaited = 0;
        }
        if (ivec != -1) {
            if (tin.v[ivec].par) {
                ixq = *reinterpret_cast<const volatile unsigned long long *>(tin.v[ivec].par);
                ipshift = (int)(ixq & 1ull) * tin.v[ivec].parshift;
            }
#if KSFD_TMAP_PARAM
            for (int sh = 0; sh < 3; ++sh)
                    asm volatile("prefetch.tensormap [%0];" ::"l"(tin.maps(ivec, 0) + sh) : "memory");
#endif
        }
        prime();
    }

    // (issuer) first loads of the chunk [k0, k1): the ring slots continue where the previous
    // chunk of this CTA (if any) stopped — every load it issued has been consumed
    __device__ __forceinline__ void prime()
    {
        if (ivec == -1) return;
        const int np = k1 - k0 + 2 * KSFD_SW;
        for (int s = 0; s < SC; ++s)
            if (s < np) issue(0, k0 - KSFD_SW + s, (cs + s) % SC);
        for (int s = 0; s < SH; ++s)
            if (s < k1 - k0) issue(1, k0 + s, (hs + s) % SH);
    }

    // (issuer) fetch plane k of this thread's input vector into ring slot `slot`:
    // halo == 0 the centre box, halo == 1 the strips
    __device__ __forceinline__ void issue(int halo, int k, int slot)
    {
        issue_vec(ivec, halo, k, slot);
    }
    __device__ __forceinline__ void issue_vec(int i, int halo, int k, int slot)
    {
        const TmaVecIn &v = tin.v[i];
        const int nc = v.nc;
        const unsigned bar = bar0 + 8 * (halo ? SC + slot : slot);
        ktma::mbar_expect(bar, (halo ? (DIM == 2 ? 4 : NH) : NTH) * 8 * nc);
        // coordinate 2 of plane k and which of the two map sets holds it
        int set = 0, kc;
        if (k < 0 || k >= g.nloc) {
            const int side = k < 0 ? 0 : 1;
            const int kk = side ? k - g.nloc : k + KSFD_SW;
            if (v.wrap) {
                kc = v.kofs[0] + (side ? kk : k + g.nloc) * nc;
            } else {
                set = 1;
                const volatile unsigned long long *f = side ? v.flag_hi : v.flag_lo;
                if (f && !(iwaited & (1u << side))) {
                    iwaited |= 1u << side;
                    tma_halo_wait(f, ixq, v.err, v.dead);
                }
                kc = v.kofs[1 + side] + ipshift + kk * nc;
            }
        } else {
            kc = v.kofs[0] + k * nc;
        }
        const CUtensorMap *m = tin.maps(i, set);
        const int i0 = tile_x0(), j0 = tile_y0();
        if (!halo) {
            ktma::load3(ktma::s32(ksfd_smem + CRING + slot * CSLOT + v.coff * NTH), m, i0, j0, kc, bar);
            return;
        }
        const int xl = i0 >= KSFD_SW ? i0 - KSFD_SW : i0 - KSFD_SW + g.n0;
        const int xr = i0 + g.ox < g.n0 ? i0 + g.ox : i0 + g.ox - g.n0;
        const double *hb = ksfd_smem + HRING + slot * HSLOT;
        if (DIM == 2) {
            ktma::load3(ktma::s32(hb + (2 * i) * L::HVEC2), m + 2, xl, 0, kc, bar);
            ktma::load3(ktma::s32(hb + (2 * i + 1) * L::HVEC2), m + 2, xr, 0, kc, bar);
        } else {
            const int yl = j0 >= KSFD_SW ? j0 - KSFD_SW : j0 - KSFD_SW + g.n1;
            const int yr = j0 + g.oy < g.n1 ? j0 + g.oy : j0 + g.oy - g.n1;
            const double *hv = hb + v.coff * NH;            // = hbox(i, 0)
            ktma::load3(ktma::s32(hv), m + 1, i0, yl, kc, bar);
            ktma::load3(ktma::s32(hv + nc * 2 * TX), m + 1, i0, yr, kc, bar);
            ktma::load3(ktma::s32(hv + nc * 4 * TX), m + 2, xl, j0, kc, bar);
            ktma::load3(ktma::s32(hv + nc * (4 * TX + 2 * TY)), m + 2, xr, j0, kc, bar);
        }
    }

    // iteration `it` of the CTA: plane kk = k0 - 2 + it enters the queue at phase PH,
    // plane kk - 2 (if owned) is emitted
    template <int PH>
    __device__ __forceinline__ void step(int it)
    {
        const int tid = threadIdx.x;
        const bool emitting = it >= 2 * KSFD_SW;        // kk - 2 >= k0: five planes are queued
        // 1. the centre box of plane kk: raw values -> pointwise fields -> queue
        ktma::mbar_wait(bar0 + 8 * cs, cph);
        if (active) {
            double cur[NPRE], f[NF];
            const int b = CRING + cs * CSLOT + tid;
#pragma unroll
            for (int c = 0; c < NPRE; ++c) cur[c] = ksfd_smem[b + c * NTH];
            op.stage(P, SmemTabs<RING>(), cur, f);
#pragma unroll
            for (int c = 0; c < NF; ++c) q[c][PH] = f[c];
        }
        const int cs_used = cs, hs_used = hs;
        if (++cs == SC) {
            cs = 0;
            cph ^= 1u;
        }
        const int ri = (it & 1) * (NF * T::SP);
        if (emitting) {
            if (Op::HAS_AUX && active) op.load_aux(g, st.e, 0, RegSink{aux});
            // 2. the halo points of the centre plane kk - 2, staged by the first warps
            if (tid < NH) {
                ktma::mbar_wait(bar0 + 8 * (SC + hs), hph);
                if (hact) {
                    double cur[NPRE], f[NF];
                    const int hb = HRING + hs * HSLOT;
#pragma unroll
                    for (int i = 0; i < NIN; ++i)
#pragma unroll
                        for (int c = 0; c < Op::nc(i); ++c)
                            cur[Op::coff(i) + c] = ksfd_smem[hb + hoff[i] + c * hfs];
                    op.stage(P, SmemTabs<RING>(), cur, f);
#pragma unroll
                    for (int c = 0; c < NF; ++c) ksfd_smem[ri + hspos + c * T::SP] = f[c];
                }
            }
            if (++hs == SH) {
                hs = 0;
                hph ^= 1u;
            }
            // 3. share the centre plane
            if (active) {
#pragma unroll
                for (int c = 0; c < NF; ++c) ksfd_smem[ri + spos + c * T::SP] = q[c][(PH + 3) % 5];
            }
        }
        __syncthreads();
        // 4. refill the slots that every thread has finished reading
        if (ivec != -1) {
            const int np = k1 - k0 + 2 * KSFD_SW;
            if (it + SC < np) issue(0, k0 - KSFD_SW + it + SC, cs_used);
            if (emitting && it - 2 * KSFD_SW + SH < k1 - k0)
                issue(1, k0 + it - 2 * KSFD_SW + SH, hs_used);
        }
        // 5. the stencil of plane kk - 2
        if (emitting) {
            if (active) {
                LaneAcc<DIM, NF, T::SP, T::SY, PH> a(q, ri + spos);
                op.emit(P, g, a, aux, st);
            }
            op.advance_out(g, st);
        }
    }

    __device__ __forceinline__ void run()
    {
        // MarchArgs::rb: the first CTA of a column marches two boundary chunks
        const int npass = (Op::NEEDS_OWNER && g.rb > 0 && blockIdx.z == 0) ? 2 : 1;
        for (int pass = 0; pass < npass; ++pass) {
            if (pass > 0) {
                if constexpr (Op::NEEDS_OWNER) {
                    march_chunk(g, blockIdx.z, k0, k1, pass);
                    op.next_chunk(g, st, k0, poff_);
                    prime();
                }
            }
            run_chunk();
            if constexpr (Op::NEEDS_OWNER) op.chunk_done(g, st, k0, k1, st.own);
        }
        bool mine = active;
        if constexpr (Op::NEEDS_OWNER) mine = st.own;
        op.finish(g, st, k0, k1, mine);
    }

    __device__ __forceinline__ void run_chunk()
    {
        const int np = k1 - k0 + 2 * KSFD_SW;
        int it = 0;
        if (!UNR) {
            for (; it < np; ++it) {
                step<4>(it);
#pragma unroll
                for (int c = 0; c < NF; ++c) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) q[c][s] = q[c][s + 1];
                }
            }
        } else {
            for (;;) {
                step<0>(it);
                if (++it >= np) break;
                step<1>(it);
                if (++it >= np) break;
                step<2>(it);
                if (++it >= np) break;
                step<3>(it);
                if (++it >= np) break;
                step<4>(it);
                if (++it >= np) break;
            }
        }
    }
};

template <int DIM, int TX, int TY, class Op, int MINB, bool UNR, int SC, int SH>
__global__ void __launch_bounds__((DIM == 2 ? TX : TX * TY), MINB)
k_tma_march(const __grid_constant__ MarchArgs g, const __grid_constant__ DevPhys P,
            const __grid_constant__ Op op, const __grid_constant__ TmaInT<Op::NIN> tin,
            const int *__restrict__ skip)
{
    KSFD_PDL_ENTER();
    if (skip && KSFD_FLAG(skip)) return;
    TmaMarcher<DIM, TX, TY, Op, UNR, SC, SH> m(g, P, op, tin);
    m.run();
}

template <int DIM, int TX, int TY, class Op, bool UNR, int SC, int SH>
constexpr size_t tma_march_smem_bytes()
{
    return sizeof(double) * TmaLayout<DIM, TX, TY, Op, SC, SH>::SMEM_DOUBLES;
}
