// Launch heuristics of the marching kernels: tile choice, balanced tile pitch,
// planes per CTA.  Included by the march_*.cu instantiation units.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>

#include "ctx.h"
#include "march_kernels.cuh"

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
static MarchPlan plan_search(const ksfd_ctx *c, const TileCand *cand, int ncand,
                             double cstage, double cemit)
{
    const Geom &g = c->g;
    MarchPlan best;
    double best_cost = 1e300;
    for (int t = 0; t < ncand; ++t) {
        if (cand[t].occ <= 0) continue;
        if (c->opt_tile_set && c->opt_tx >= 0 && c->opt_tx < ncand && t != c->opt_tx)
            continue;
        const int ntx = (g.n0 + cand[t].TX - 1) / cand[t].TX;
        const int ox = (g.n0 + ntx - 1) / ntx;
        const int nty = c->dim == 3 ? (g.n1 + cand[t].TY - 1) / cand[t].TY : 1;
        const int oy = c->dim == 3 ? (g.n1 + nty - 1) / nty : 1;
        const long long cols = (long long)ntx * nty;
        // active warps of the stage / emit phases
        double wstage, wemit;
        if (c->dim == 2) {
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
                            double cstage, double cemit)
{
    if (!c->plan_cache) c->plan_cache = new PlanMap();
    PlanMap &pm = *static_cast<PlanMap *>(c->plan_cache);
    auto it = pm.find(key);
    if (it != pm.end()) return it->second;
    MarchPlan p = plan_search(c, cand, ncand, cstage, cemit);
    pm[key] = p;
    return p;
}

template <int DIM, int TX, int TY, class Op, int MINB, bool UNR, int DEPTH>
static int tile_occupancy()
{
    static int occ = -1;
    if (occ >= 0) return occ;
    using T = TileT<DIM, TX, TY>;
    auto kern = k_march<DIM, TX, TY, Op, MINB, UNR, DEPTH>;
    const size_t smem = march_smem_bytes<Op, T::SP, T::NT, DEPTH>();
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess) {
        cudaGetLastError();
        occ = 0;
        return occ;
    }
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, T::NT, smem) !=
        cudaSuccess) {
        cudaGetLastError();
        nb = 0;
    }
    occ = nb;
    return occ;
}

template <int DIM, int TX, int TY, class Op, int MINB, bool UNR, int DEPTH>
static int launch_tile(const ksfd_ctx *c, const Op &op, const MarchPlan &p, const int *skip,
                       cudaStream_t st)
{
    using T = TileT<DIM, TX, TY>;
    auto kern = k_march<DIM, TX, TY, Op, MINB, UNR, DEPTH>;
    const size_t smem = march_smem_bytes<Op, T::SP, T::NT, DEPTH>();
    KSFD_KLAUNCH(kern, p.grid, T::NT, smem, st, p.a, c->P, op, skip);
    CKL();
    return 0;
}

// two tile candidates per operator: (AX, AY, AMINB) and (BX, BY, BMINB)
template <int DIM, class Op, bool UNR, int DEPTH, int AX, int AY, int AMINB, int BX, int BY,
          int BMINB>
static int launch_op(ksfd_ctx *c, const Op &op, int opkey, double cstage, double cemit,
                     const int *skip, cudaStream_t st)
{
    const long long maxel = (long long)(c->g.nloc + 2 * KSFD_SW) * c->g.plane_pts * (c->dof + 2);
    if (maxel >= (1LL << 31))
        return fail("rank-local slab too large for the 32-bit indexed kernels; "
                    "decompose over more ranks");
    TileCand cand[2] = {
        {AX, AY, TileT<DIM, AX, AY>::NT, tile_occupancy<DIM, AX, AY, Op, AMINB, UNR, DEPTH>()},
        {BX, BY, TileT<DIM, BX, BY>::NT, tile_occupancy<DIM, BX, BY, Op, BMINB, UNR, DEPTH>()}};
    if (cand[0].occ == 0 && cand[1].occ == 0)
        return fail("marching kernel does not fit on this device");
    MarchPlan p = plan_march(c, opkey * 100 + DIM * 10 + Op::NF, cand, 2, cstage, cemit);
    if (p.tile < 0) return fail("no marching tile fits");
    if (p.tile == 0) return launch_tile<DIM, AX, AY, Op, AMINB, UNR, DEPTH>(c, op, p, skip, st);
    return launch_tile<DIM, BX, BY, Op, BMINB, UNR, DEPTH>(c, op, p, skip, st);
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
