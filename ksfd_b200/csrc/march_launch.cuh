// Launch heuristics of the marching kernels: tile choice, balanced tile pitch,
// planes per CTA.  Included by the march_*.cu instantiation units.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>

#include <array>
#include <tuple>

#include "ctx.h"
#include "march_kernels.cuh"
#include "tma_host.h"
#include "tma_march.cuh"

struct MarchPlan {
    int tile = -1;              // index into the candidate list
    MarchArgs a{};
    dim3 grid;
};
struct TileCand {
    int TX, TY, NT, occ;        // max outputs per tile, threads, resident CTAs/SM
};
typedef std::map<long long, MarchPlan> PlanMap;

// Cost model: waves of CTAs over the SMs times the work of one CTA (stage cost
// on every active lane of rz+4 planes + emit cost on the outputs of rz planes +
// a fixed start-up cost); work is counted in whole warps.
// tma: tiles of the TMA-fed marcher (tma_march.cuh): even pitches, one thread per output,
// the halo points staged in an extra pass of the first warps
static MarchPlan plan_search(const ksfd_ctx *c, const TileCand *cand, int ncand,
                             double cstage, double cemit, bool tma = false)
{
    const Geom &g = c->g;
    MarchPlan best;
    double best_cost = 1e300;
    for (int t = 0; t < ncand; ++t) {
        if (cand[t].occ <= 0) continue;
        if (c->opt_tile_set && c->opt_tx >= 0 && c->opt_tx < ncand && t != c->opt_tx)
            continue;
        const int ntx = (g.n0 + cand[t].TX - 1) / cand[t].TX;
        int ox = (g.n0 + ntx - 1) / ntx;
        const int nty = c->dim == 3 ? (g.n1 + cand[t].TY - 1) / cand[t].TY : 1;
        int oy = c->dim == 3 ? (g.n1 + nty - 1) / nty : 1;
        if (tma) {
            ox += ox & 1;
            if (c->dim == 3) oy += oy & 1;
        }
        const long long cols = (long long)ntx * nty;
        // active warps of the stage / emit phases
        double wstage, wemit;
        if (tma) {
            wemit = cand[t].NT / 32.0;
            wstage = wemit + (c->dim == 2 ? 1.0 : std::ceil((4.0 * cand[t].TX + 4.0 * cand[t].TY) / 32.0));
        } else if (c->dim == 2) {
            wstage = std::ceil((ox + 4) / 32.0);
            wemit = wstage;
        } else {
            // interior rows are TX wide: a row of ox<TX outputs still occupies
            // its warps
            wemit = std::ceil(cand[t].TX * oy / 32.0);
            wstage = wemit + std::ceil(4.0 * cand[t].TX / 32.0) + std::ceil(4.0 * oy / 32.0);
        }
        const int occ = cand[t].occ;
        const double slots = (double)c->sm_count * occ;
        for (int chunks = 1; chunks <= g.nloc;
             chunks = chunks < 32 ? chunks + 1 : chunks + chunks / 16) {
            int rz = (g.nloc + chunks - 1) / chunks;
            if (c->opt_rz > 0) rz = std::min(c->opt_rz, g.nloc);
            const int nch = (g.nloc + rz - 1) / rz;
            if (rz < 2 && g.nloc > 2) break;
            const double ctas = (double)cols * nch;
            const double per_cta = wstage * (rz + 2 * KSFD_SW) * cstage + wemit * rz * cemit +
                                   100.0 * wstage;
            // CTAs run `occ` at a time per SM and share its pipes: time ~ work
            // per SM, rounded up to whole waves of resident CTAs
            const double waves = std::ceil(ctas / slots);
            const double cost = waves * occ * per_cta;
            if (cost < best_cost) {
                best_cost = cost;
                best.tile = t;
                best.a.n0 = g.n0;
                best.a.n1 = g.n1;
                best.a.nloc = g.nloc;
                best.a.fs = (int)g.plane_pts;
                best.a.ox = ox;
                best.a.oy = oy;
                best.a.rz = rz;
                best.grid = dim3(ntx, nty, nch);
            }
            if (c->opt_rz > 0) break;
        }
    }
    return best;
}

static MarchPlan plan_march(ksfd_ctx *c, long long key, const TileCand *cand, int ncand,
                            double cstage, double cemit, bool tma = false)
{
    if (!c->plan_cache) c->plan_cache = new PlanMap();
    PlanMap &pm = *static_cast<PlanMap *>(c->plan_cache);
    auto it = pm.find(key);
    if (it != pm.end()) return it->second;
    MarchPlan p = plan_search(c, cand, ncand, cstage, cemit, tma);
    pm[key] = p;
    return p;
}

template <int DIM, int TX, int TY, class Op, int MINB, bool UNR>
static int tile_occupancy(int device)
{
    // per device: the dynamic-shared-memory opt-in below is a per-device attribute
    static std::map<int, int> occ_of;
    auto it = occ_of.find(device);
    if (it != occ_of.end()) return it->second;
    int occ;
    using T = TileT<DIM, TX, TY>;
    auto kern = k_march<DIM, TX, TY, Op, MINB, UNR>;
    const size_t smem = march_smem_bytes<Op, T::SP>();
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess) {
        cudaGetLastError();
        occ_of[device] = 0;
        return 0;
    }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, T::NT, smem) !=
        cudaSuccess) {
        cudaGetLastError();
        nb = 0;
    }
    occ = nb;
    occ_of[device] = occ;
    return occ;
}

template <int DIM, int TX, int TY, class Op, int MINB, bool UNR>
static int launch_tile(const ksfd_ctx *c, const Op &op, const MarchPlan &p, const int *skip,
                       cudaStream_t st)
{
    using T = TileT<DIM, TX, TY>;
    auto kern = k_march<DIM, TX, TY, Op, MINB, UNR>;
    const size_t smem = march_smem_bytes<Op, T::SP>();
    // Jacobian-side operators run on the physics of the linearisation (ctx.h: Pjac)
    const DevPhys &P = Op::JACOBIAN ? c->Pjac : c->P;
    KSFD_KLAUNCH(kern, p.grid, T::NT, smem, st, p.a, P, op, skip);
    CKL();
    return 0;
}

// two tile candidates per operator: (AX, AY, AMINB) and (BX, BY, BMINB)
template <int DIM, class Op, bool UNR, int AX, int AY, int AMINB, int BX, int BY, int BMINB>
static int launch_op(ksfd_ctx *c, const Op &op, int opkey, double cstage, double cemit,
                     const int *skip, cudaStream_t st, int max_ctas = 0)
{
    const long long maxel = (long long)(c->g.nloc + 2 * KSFD_SW) * c->g.plane_pts * (c->dof + 2);
    if (maxel >= (1LL << 31))
        return fail("rank-local slab too large for the 32-bit indexed kernels; "
                    "decompose over more ranks");
    TileCand cand[2] = {
        {AX, AY, TileT<DIM, AX, AY>::NT, tile_occupancy<DIM, AX, AY, Op, AMINB, UNR>(c->device)},
        {BX, BY, TileT<DIM, BX, BY>::NT, tile_occupancy<DIM, BX, BY, Op, BMINB, UNR>(c->device)}};
    if (cand[0].occ == 0 && cand[1].occ == 0)
        return fail("marching kernel does not fit on this device");
    MarchPlan p = plan_march(c, opkey * 100 + DIM * 10 + Op::NF, cand, 2, cstage, cemit);
    if (p.tile < 0) return fail("no marching tile fits");
    if (max_ctas && (long long)p.grid.x * p.grid.y * p.grid.z > max_ctas)
        return fail("marching grid exceeds the per-CTA reduction buffer");
    if (p.tile == 0) return launch_tile<DIM, AX, AY, Op, AMINB, UNR>(c, op, p, skip, st);
    return launch_tile<DIM, BX, BY, Op, BMINB, UNR>(c, op, p, skip, st);
}

// ---------------------------------------------------------------------------
// TMA-fed marcher (tma_march.cuh)
// ---------------------------------------------------------------------------
// encoded tensor maps are cached per (buffer, extent, fields per plane, tile)
typedef std::tuple<const void *, long long, int, int, int> TmapKey;
typedef std::map<TmapKey, std::array<CUtensorMap, 3>> TmapCache;

// option `variant`: 0 auto, 1 direct kernels, 2 marching (TMA-fed where eligible),
// 3 marching with the register-prefetch kernels (k_march) only
static bool ksfd_use_tma(const ksfd_ctx *c)
{
    if (c->variant == 3) return false;
    // rows must be multiples of 16 bytes (tensor-map strides) and tile pitches even
    if (c->g.n0 % 2 != 0 || (c->dim == 3 && c->g.n1 % 2 != 0)) return false;
    return ksfd_tmap_encoder() != nullptr;
}

static int tma_maps(ksfd_ctx *c, CUtensorMap out[3], const double *buf, long long fields, int nc,
                    int TX, int TY)
{
    if (!c->tmap_cache) c->tmap_cache = new TmapCache();
    TmapCache &tc = *static_cast<TmapCache *>(c->tmap_cache);
    const TmapKey key(buf, fields, nc, TX, TY);
    auto it = tc.find(key);
    if (it == tc.end()) {
        std::array<CUtensorMap, 3> m;
        const std::string e = ksfd_make_tmaps(m.data(), buf, c->g.n0, c->g.n1, fields, nc, TX, TY);
        if (!e.empty()) return fail("tensor map: " + e);
        if (tc.size() > 4096) tc.clear();
        it = tc.emplace(key, m).first;
    }
    for (int s = 0; s < 3; ++s) out[s] = it->second[s];
    return 0;
}

template <class Op>
static int tma_bind(ksfd_ctx *c, TmaInT<Op::NIN> &tin, const TmaSrc *src, int TX, int TY)
{
    for (int i = 0; i < Op::NIN; ++i) {
        const TmaSrc &t = src[i];
        const int nc = Op::nc(i);
        TRY(tma_maps(c, tin.m[i][0], t.base, t.base_fields, nc, TX, TY));
        if (t.halo)
            TRY(tma_maps(c, tin.m[i][1], t.halo, t.halo_fields, nc, TX, TY));
        else
            for (int s = 0; s < 3; ++s) tin.m[i][1][s] = tin.m[i][0][s];
        tin.v[i].kofs[0] = t.k0;
        tin.v[i].kofs[1] = t.klo;
        tin.v[i].kofs[2] = t.khi;
        tin.v[i].wrap = t.wrap;
        tin.v[i].par = t.par;
        tin.v[i].parshift = t.parshift;
        tin.v[i].pad_ = 0;
        tin.v[i].nc = nc;
        tin.v[i].coff = Op::coff(i);
        tin.v[i].flag_lo = t.flag_lo;
        tin.v[i].flag_hi = t.flag_hi;
        tin.v[i].err = t.err;
        tin.v[i].dead = t.dead;
    }
    return 0;
}

template <int DIM, int TX, int TY, class Op, int MINB, bool UNR, int SC, int SH>
static int tma_tile_occupancy(const ksfd_ctx *c)
{
    static std::map<int, int> occ_of;           // per device (the opt-in is per device too)
    auto it = occ_of.find(c->device);
    if (it != occ_of.end()) return it->second;
    using M = TmaMarcher<DIM, TX, TY, Op, UNR, SC, SH>;
    auto kern = k_tma_march<DIM, TX, TY, Op, MINB, UNR, SC, SH>;
    const size_t smem = tma_march_smem_bytes<DIM, TX, TY, Op, UNR, SC, SH>();
    int nb = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, M::NTH, smem) != cudaSuccess) {
        cudaGetLastError();
        nb = 0;
    }
    occ_of[c->device] = nb;
    return nb;
}

template <int DIM, int TX, int TY, class Op, int MINB, bool UNR, int SC, int SH>
static int launch_tma_tile(ksfd_ctx *c, const Op &op, const TmaSrc *src, const MarchPlan &p,
                           const int *skip, cudaStream_t st)
{
    using M = TmaMarcher<DIM, TX, TY, Op, UNR, SC, SH>;
    auto kern = k_tma_march<DIM, TX, TY, Op, MINB, UNR, SC, SH>;
    const size_t smem = tma_march_smem_bytes<DIM, TX, TY, Op, UNR, SC, SH>();
    TmaInT<Op::NIN> tin;
    TRY(tma_bind<Op>(c, tin, src, TX, TY));
    const DevPhys &P = Op::JACOBIAN ? c->Pjac : c->P;
    KSFD_KLAUNCH(kern, p.grid, M::NTH, smem, st, p.a, P, op, tin, skip);
    CKL();
    return 0;
}

// two tile candidates per operator: (AX, AY, AMINB, ASC) and (BX, BY, BMINB, BSC)
template <int DIM, class Op, bool UNR, int AX, int AY, int AMINB, int ASC, int BX, int BY,
          int BMINB, int BSC>
static int launch_tma_op(ksfd_ctx *c, const Op &op, const TmaSrc *src, int opkey, double cstage,
                         double cemit, const int *skip, cudaStream_t st, int max_ctas = 0,
                         bool bnd_first = false)
{
    const long long maxel = (long long)(c->g.nloc + 2 * KSFD_SW) * c->g.plane_pts * (c->dof + 2);
    if (maxel >= (1LL << 31))
        return fail("rank-local slab too large for the 32-bit indexed kernels; "
                    "decompose over more ranks");
    constexpr int ANT = DIM == 2 ? AX : AX * AY, BNT = DIM == 2 ? BX : BX * BY;
    TileCand cand[2] = {
        {AX, AY, ANT, tma_tile_occupancy<DIM, AX, AY, Op, AMINB, UNR, ASC, ASC>(c)},
        {BX, BY, BNT, tma_tile_occupancy<DIM, BX, BY, Op, BMINB, UNR, BSC, BSC>(c)}};
    if (cand[0].occ == 0 && cand[1].occ == 0)
        return fail("TMA marching kernel does not fit on this device");
    MarchPlan p = plan_march(c, 1000 + opkey * 100 + DIM * 10 + Op::NF, cand, 2, cstage, cemit, true);
    if (p.tile < 0) return fail("no marching tile fits");
    if (max_ctas && (long long)p.grid.x * p.grid.y * p.grid.z > max_ctas)
        return fail("marching grid exceeds the per-CTA reduction buffer");
    if (bnd_first && p.grid.z >= 3) {
        // the boundary planes in the first CTA of every column, as two short chunks
        // (MarchArgs::rb): same number of CTAs, the other chunks keep rz planes
        const int nch = (int)p.grid.z, rz = p.a.rz, nloc = p.a.nloc;
        int rb = (nloc - (nch - 1) * rz + 1) / 2;
        rb = std::max(KSFD_SW, std::min(rb, rz / 2));
        const int mid = nloc - 2 * rb;
        if (mid <= (nch - 1) * rz && mid > (nch - 2) * rz && rb >= KSFD_SW) p.a.rb = rb;
    }
    if (p.tile == 0)
        return launch_tma_tile<DIM, AX, AY, Op, AMINB, UNR, ASC, ASC>(c, op, src, p, skip, st);
    return launch_tma_tile<DIM, BX, BY, Op, BMINB, UNR, BSC, BSC>(c, op, src, p, skip, st);
}

#define KSFD_DISPATCH_NLIG(FN, ...)                                        \
    do {                                                                   \
        switch (c->dof - 1) {                                              \
        case 1: return FN<1>(__VA_ARGS__);                                 \
        case 2: return FN<2>(__VA_ARGS__);                                 \
        case 3: return FN<3>(__VA_ARGS__);                                 \
        case 4: return FN<4>(__VA_ARGS__);                                 \
        }                                                                  \
        return fail("no marching kernel for this dof");                    \
    } while (0)

#define KSFD_CAT_(a, b) a##b
#define KSFD_CAT(a, b) KSFD_CAT_(a, b)
