// ksfd_b200: host side of the C ABI (include/ksfd_b200.h): context, kernel
// launch heuristics, NCCL halo ring, device-resident GMRES, ROSW/BEuler step.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <dlfcn.h>
#include <map>
#include <atomic>
#include <string>
#include <vector>

#include "blas1_kernels.cuh"
#include "sweep_op.cuh"
#include "ctx.h"
#include "march_launch.cuh"
#include "naive_kernels.cuh"
#include "fftpc.cuh"

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

int ksfd_fail(const std::string &m)
{
    g_err = m;
    return 1;
}
void ksfd_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------
// NCCL through dlopen (torch ships libnccl.so.2; no header needed)
// ---------------------------------------------------------------------------
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSum_ = 0, ncclMax_ = 2 };
enum { ncclFloat64_ = 8 };
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t,
                     cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load(const char *path)
{
    if (g_nccl.h) return 0;
    void *h = nullptr;
    if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(std::string("cannot dlopen libnccl: ") + dlerror());
    g_nccl.h = h;
#define SYM(n)                                                     \
    *(void **)(&g_nccl.n) = dlsym(h, "nccl" #n);                   \
    if (!g_nccl.n) return fail("libnccl lacks symbol nccl" #n);
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(Send) SYM(Recv)
    SYM(AllReduce) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString)
#undef SYM
    return 0;
}
#define NK(call)                                                          \
    do {                                                                  \
        int r_ = (call);                                                  \
        if (r_ != 0)                                                      \
            return fail(std::string(#call) + ": " + g_nccl.GetErrorString(r_)); \
    } while (0)

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
// largest GMRES restart length: the Arnoldi column + its extras must fit one
// peer-to-peer reduce contribution (KSFD_P2P_RED_MAX) and the shared-memory
// Hessenberg of k_gm_finalize (KSFD_GM_MAXM)
#define KSFD_MAX_RESTART 60
#define SC_USER 440              // ksfd_norm2 / ksfd_sum_dof0 results (clear of the solver's slots)
#define SC_H 0                  // Hessenberg column (<= 128)
#define SC_H2 128               // second Gram-Schmidt pass
#define SC_NORM 300
#define SC_ENORM 301
#define SC_AUX 126               // [inv h_{j+1}, cancellation flag], right after the column
static long long nlocal(const ksfd_ctx *c) { return c->g.npts * c->dof; }

// in-situ kernel timing: CUDA event pairs around launches (kind 0 J.v stencil, 1 residual
// stencil, 2 multi-dot (+ rank sum, Givens), 3 orthogonalise-and-scale (+ halo push),
// 4 first Krylov vector, 5 start of a cycle (norm / true residual), 6 Richardson sweep,
// 7 unused); scopes nest
struct ProfRec {
    int kind, start, stop;
};
struct ProfState {
    std::vector<cudaEvent_t> pool;          // reused events
    std::vector<ProfRec> recs;
    size_t used = 0;
};
static int prof_event(ProfState *ps, cudaStream_t st)
{
    if (ps->used == ps->pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ps->pool.push_back(e);
    }
    cudaEventRecord(ps->pool[ps->used], st);
    return (int)ps->used++;
}
struct ProfScope {
    ksfd_ctx *c;
    cudaStream_t st;
    int kind, start;
    ProfScope(ksfd_ctx *c_, int kind_, cudaStream_t st_)
        : c(c_->prof_on ? c_ : nullptr), st(st_), kind(kind_), start(-1)
    {
        if (!c) return;
        if (!c->prof) c->prof = new ProfState();
        ProfState *ps = static_cast<ProfState *>(c->prof);
        if (ps->used > 200000) {        // a forgotten profile switch must not grow without bound
            c = nullptr;
            return;
        }
        start = prof_event(ps, st);
    }
    ~ProfScope()
    {
        if (!c) return;
        ProfState *ps = static_cast<ProfState *>(c->prof);
        const int stop = prof_event(ps, st);
        ps->recs.push_back(ProfRec{kind, start, stop});
    }
};

static int ensure_work(ksfd_ctx *c, int i)
{
    if (!c->work[i]) CK(cudaMalloc(&c->work[i], sizeof(double) * nlocal(c)));
    return 0;
}

static void fftpc_destroy(ksfd_ctx *c);

extern "C" int ksfd_abi_version(void) { return KSFD_ABI_VERSION; }
extern "C" const char *ksfd_last_error(void) { return g_err.c_str(); }
extern "C" int64_t ksfd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int ksfd_ctx_create(ksfd_ctx **out, int dim, const int64_t n_global[3],
                               int64_t last_start, int64_t last_count, int dof,
                               int device)
{
    if (!out) return fail("ksfd_ctx_create: out is NULL");
    if (dim < 1 || dim > 3) return fail("KSFD.Grid dimension must be 1, 2, or 3");
    if (dof < 2 || dof > KSFD_MAX_LIGANDS + 1)
        return fail("dof must be in [2, " + std::to_string(KSFD_MAX_LIGANDS + 1) + "]");
    for (int d = 0; d < dim; ++d)
        if (n_global[d] < 1) return fail("grid extents must be >= 1");
    if (last_count < KSFD_SW)
        return fail("each rank must own at least stencil_width planes");
    if (last_start < 0 || last_start + last_count > n_global[dim - 1])
        return fail("ownership range outside the grid");
    CK(cudaSetDevice(device));
    ksfd_ctx *c = new ksfd_ctx();
    c->dim = dim;
    c->dof = dof;
    c->device = device;
    for (int d = 0; d < dim; ++d) c->n[d] = n_global[d];
    c->last_start = last_start;
    c->last_count = last_count;
    c->last_global = n_global[dim - 1];
    Geom &g = c->g;
    g.dim = dim;
    g.dof = dof;
    g.n0 = dim >= 2 ? (int)c->n[0] : 1;
    g.n1 = dim >= 3 ? (int)c->n[1] : 1;
    g.nloc = (int)last_count;
    g.plane_pts = (long long)g.n0 * g.n1;
    g.npts = g.plane_pts * g.nloc;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->max_smem = (int)prop.sharedMemPerBlockOptin;
    CK(cudaMalloc(&c->partial, sizeof(double) * (KSFD_MAXV + 1) * KSFD_RED_BLOCKS));
    CK(cudaMalloc(&c->dscal, sizeof(double) * KSFD_NSCAL));
    CK(cudaMallocHost(&c->hscal, sizeof(double) * KSFD_NSCAL));
    c->halo_plane_doubles = (size_t)g.plane_pts * (dof + 2);
    if (const char *e = getenv("KSFD_GM_RUNAHEAD")) c->gm_runahead = std::max(0, std::min(atoi(e), 8));
    *out = c;
    return 0;
}

extern "C" int ksfd_ctx_destroy(ksfd_ctx *c)
{
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    for (auto &h : c->halo) cudaFree(h);
    cudaFree(c->coef);
    cudaFree(c->pc);
    cudaFree(c->partial);
    cudaFree(c->dscal);
    cudaFreeHost(c->hscal);
    cudaFree(c->krylov);
    fftpc_destroy(c);
    cudaFree(c->fft_spec);
    cudaFree(c->fft_spec2);
    cudaFree(c->fft_means);
    cudaFree(c->gm);
    cudaFree(c->sw_partial);
    cudaFree(c->gmi);
    cudaFree(c->gm_done);
    for (int r = 0; r < 16; ++r)
        if (c->p2p_peer[r] && c->p2p_peer[r] != c->p2p_mine) cudaIpcCloseMemHandle(c->p2p_peer[r]);
    cudaFree(c->p2p_mine);
    cudaFree(c->p2p_done);
    cudaFree(c->p2p_ctr);
    if (c->p2p_err) cudaFreeHost(c->p2p_err);
    if (c->gm_status) cudaFreeHost(c->gm_status);
    for (auto &w : c->work) cudaFree(w);
    if (c->prof) {
        ProfState *ps = static_cast<ProfState *>(c->prof);
        for (cudaEvent_t e : ps->pool) cudaEventDestroy(e);
        delete ps;
    }
    ksfd_free_plans(c);
    delete c;
    return 0;
}

extern "C" int64_t ksfd_local_size(const ksfd_ctx *c) { return c ? nlocal(c) : 0; }

#define KSFD_PROF_KINDS 8
extern "C" int ksfd_profile_fetch(ksfd_ctx *c, double out[4 * KSFD_PROF_KINDS], void *stream)
{
    if (!c || !out) return fail("ksfd_profile_fetch: NULL argument");
    for (int i = 0; i < 4 * KSFD_PROF_KINDS; ++i) out[i] = 0.0;
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    ProfState *ps = static_cast<ProfState *>(c->prof);
    if (!ps) return 0;
    std::vector<float> ms(ps->recs.size(), 0.f);
    float mx[KSFD_PROF_KINDS] = {0.f};
    for (size_t i = 0; i < ps->recs.size(); ++i) {
        const ProfRec &o = ps->recs[i];
        if (cudaEventElapsedTime(&ms[i], ps->pool[o.start], ps->pool[o.stop]) != cudaSuccess) {
            cudaGetLastError();
            ms[i] = 0.f;
        }
        mx[o.kind] = std::max(mx[o.kind], ms[i]);
    }
    // launches of the pipelined solver that were made ahead of a convergence test and
    // returned at once (~2 us) are not real passes: a launch is ACTIVE when it lasted at
    // least 4 us
    for (size_t i = 0; i < ps->recs.size(); ++i) {
        const int k = ps->recs[i].kind;
        out[4 * k + 2] += 1.0;
        out[4 * k + 3] += ms[i];
        if (ms[i] >= 0.004f) {
            out[4 * k] += 1.0;
            out[4 * k + 1] += ms[i];
        }
    }
    ps->recs.clear();
    ps->used = 0;
    return 0;
}

extern "C" int ksfd_set_physics(ksfd_ctx *c, const ksfd_physics *p)
{
    if (!c || !p) return fail("ksfd_set_physics: NULL argument");
    if (p->nlig != c->dof - 1)
        return fail("physics.nlig must equal dof-1");
    if (p->ngroups < 0 || p->ngroups > KSFD_MAX_GROUPS)
        return fail("too many ligand groups");
    if (!(p->cushion != 0.0) || !(p->rhomax != 0.0))
        return fail("cushion and rhomax must be non-zero");
    DevPhys &P = c->P;
    P.ngroups = p->ngroups;
    P.nlig = p->nlig;
    P.cap_type = p->cap_type;
    P.dim = c->dim;
    P.s2 = p->s2;
    P.rhomax = p->rhomax;
    P.inv_cushion = 1.0 / p->cushion;
    P.capscale = p->maxscale * p->s2;
    P.rhomin = p->rhomin;
    P.Umin = p->Umin;
    P.inv_rhomax = 1.0 / p->rhomax;
    // group slots beyond ngroups are neutral (alpha 1, beta 0): the marching
    // kernels evaluate all slots without tests
    for (int g = 0; g < KSFD_MAX_GROUPS; ++g) {
        P.alpha[g] = g < p->ngroups ? p->alpha[g] : 1.0;
        P.beta[g] = g < p->ngroups ? p->beta[g] : 0.0;
    }
    for (int l = 0; l < KSFD_MAX_LIGANDS; ++l) {
        P.lig_group[l] = l < p->nlig ? p->lig_group[l] : -1;
        P.weight[l] = p->weight[l];
        P.s[l] = p->s[l];
        P.gamma[l] = p->gamma[l];
        P.D[l] = p->D[l];
        if (l < p->nlig && (p->lig_group[l] < 0 || p->lig_group[l] >= p->ngroups))
            return fail("lig_group out of range");
    }
    P.lig_group[KSFD_MAX_LIGANDS] = -1;
    // the reference lays stencil axes out as x, y, z = axis 0,1,2; the last
    // axis is the marching / decomposed one in every dimension.
    for (int a = 0; a < 3; ++a)
        for (int s = 0; s < 5; ++s) {
            P.w1[a][s] = a < c->dim ? p->w1[a][s] : 0.0;
            P.w2[a][s] = a < c->dim ? p->w2[a][s] : 0.0;
        }
    for (int g = 0; g < KSFD_MAX_GROUPS; ++g)
        for (int l = 0; l < KSFD_MAX_LIGANDS; ++l)
            P.Wgl[g][l] = (l < p->nlig && p->lig_group[l] == g) ? p->weight[l] : 0.0;
    P.w2c = 0.0;
    for (int a = 0; a < c->dim; ++a) P.w2c += P.w2[a][2];
    // symmetric-stencil constants and the check that the reference weights
    // follow the (1,-8,0,8,-1)/(12h), (-1,16,-30,16,-1)/(12h^2) pattern to
    // rounding (they do for every order-3 grid; KSFD/ksfdsym.py:391-436)
    P.sym_ok = 1;
    for (int a = 0; a < 3; ++a) {
        P.c1[a] = P.c2[a] = P.c1sq[a] = 0.0;
        if (a >= c->dim) continue;
        const double c1 = P.w1[a][3] / 8.0, c2 = P.w2[a][1] / 16.0;
        P.c1[a] = c1;
        P.c2[a] = c2;
        P.c1sq[a] = c1 * c1;
        const double e1 = 8e-16 * std::fabs(8.0 * c1), e2 = 8e-16 * std::fabs(30.0 * c2);
        const double r1[5] = {c1, -8.0 * c1, 0.0, 8.0 * c1, -c1};
        const double r2[5] = {-c2, 16.0 * c2, -30.0 * c2, 16.0 * c2, -c2};
        for (int s = 0; s < 5; ++s)
            if (!(std::fabs(P.w1[a][s] - r1[s]) <= e1) || !(std::fabs(P.w2[a][s] - r2[s]) <= e2))
                P.sym_ok = 0;
    }
    P.ycap1 = -2.0 * P.inv_cushion;
    P.ycap0 = 2.0 * P.rhomax * P.inv_cushion;
    P.capscale2 = 2.0 * P.capscale;
    P.mk = fastk_default();
    c->have_phys = true;
    return 0;
}

extern "C" int ksfd_set_option(ksfd_ctx *c, const char *key, int64_t v)
{
    if (!c || !key) return fail("ksfd_set_option: NULL argument");
    std::string k(key);
    if (k == "variant") c->variant = (int)v;
    else if (k == "tile") { c->opt_tx = (int)v; c->opt_tile_set = v >= 0; }
    else if (k == "rz") c->opt_rz = (int)v;
    else if (k == "profile") c->prof_on = v != 0;
    else if (k == "gmres_pipeline") c->gm_pipeline = (int)v;
    else if (k == "halo_p2p") c->p2p_on = v != 0 && c->p2p_up != nullptr;   // same on all ranks
    else if (k == "gmres_cycle_exp") c->gm_cycle_factor = v <= 0 ? 0.0 : std::pow(10.0, -(double)v);
    else if (k == "sweep_test_lead") c->sw_test_lead = (int)std::max<int64_t>(1, v);
    else if (k == "sweep_fuse_push") c->sw_fuse_push = v != 0;
    else if (k == "fuse_push_mask") c->fuse_push_mask = (int)v;
    else if (k == "speculate") c->spec_on = v != 0;
    else if (k == "sweep_underpredict") c->sw_underpredict = (int)std::max<int64_t>(0, std::min<int64_t>(v, 4));
    else if (k == "gmres_runahead") c->gm_runahead = (int)std::max<int64_t>(0, std::min<int64_t>(v, 8));
    else return fail("unknown option " + k);
    ksfd_invalidate_plans(c);
    return 0;
}

// ---------------------------------------------------------------------------
// halos
// ---------------------------------------------------------------------------
extern "C" int ksfd_nccl_unique_id(const char *path, char id_out[128])
{
    TRY(nccl_load(path));
    ncclUniqueId id;
    NK(g_nccl.GetUniqueId(&id));
    memcpy(id_out, id.internal, 128);
    return 0;
}

extern "C" int ksfd_comm_init(ksfd_ctx *c, const char *path, int nranks, int rank,
                              const char id[128])
{
    if (!c) return fail("ksfd_comm_init: NULL ctx");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail("bad rank/nranks");
    c->nranks = nranks;
    c->rank = rank;
    if (nranks == 1) return 0;
    TRY(nccl_load(path));
    CK(cudaSetDevice(c->device));
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    NK(g_nccl.CommInitRank(&c->comm, nranks, uid, rank));
    return 0;
}

// VecRef for a vector with `stride` doubles per point.  One rank: the ghost
// planes alias the vector (periodic wrap).  Several ranks: ghost planes live in
// halo slot `slot`, filled by exchange().
static size_t p2p_buf_doubles(const ksfd_ctx *c);
static double *p2p_buf(const ksfd_ctx *c, double *base, int slot, int parity);

static VecRef make_ref(const ksfd_ctx *c, const double *base, int stride, int slot)
{
    VecRef r;
    r.base = base;
    r.par = nullptr;
    r.pstride = 0;
    r.flag_lo = r.flag_hi = nullptr;
    r.err = nullptr;
    r.dead = nullptr;
    const long long ps = c->g.plane_pts * stride;
    if (c->nranks == 1) {
        r.lo = base + (long long)(c->g.nloc - KSFD_SW) * ps;
        r.hi = base;
    } else if (c->p2p_on) {
        // parity-0 buffer; the kernel adds (*par & 1) * pstride
        r.lo = p2p_buf(c, c->p2p_mine, slot, 0);
        r.hi = r.lo + KSFD_SW * c->halo_plane_doubles;
        r.par = c->p2p_ctr + slot;
        r.pstride = (long long)p2p_buf_doubles(c);
        // the marching kernels wait for the neighbours' flags themselves (already
        // satisfied when the exchange was made by k_halo_xchg, which waits)
        typedef const volatile unsigned long long *cflag_t;
        r.flag_lo = reinterpret_cast<cflag_t>(c->p2p_mine) + slot * 2 + 0;
        r.flag_hi = reinterpret_cast<cflag_t>(c->p2p_mine) + slot * 2 + 1;
        r.err = c->p2p_err_dev;
        r.dead = c->p2p_ctr + KSFD_HALO_SLOTS + 1;
    } else {
        r.lo = c->halo[slot];
        r.hi = c->halo[slot] + KSFD_SW * c->halo_plane_doubles;
    }
    return r;
}

// VecRef + the buffer description the TMA-fed marcher needs (ctx.h: TmaSrc)
static HostVec make_hvec(const ksfd_ctx *c, const double *base, int stride, int slot)
{
    HostVec h;
    h.r = make_ref(c, base, stride, slot);
    TmaSrc &t = h.t;
    t.base = base;
    t.base_fields = (long long)c->g.nloc * stride;
    t.k0 = 0;
    const long long hp = (long long)(c->halo_plane_doubles / c->g.plane_pts);   // fields per ghost plane
    if (c->nranks == 1) {
        t.wrap = 1;
    } else if (c->p2p_on) {
        // [parity 0: lo | hi][parity 1: lo | hi], planes packed with the vector's stride
        t.wrap = 0;
        t.halo = p2p_buf(c, c->p2p_mine, slot, 0);
        t.halo_fields = 2 * 2 * KSFD_SW * hp;
        t.klo = 0;
        t.khi = (int)(KSFD_SW * hp);
        t.par = c->p2p_ctr + slot;
        t.parshift = (int)(2 * KSFD_SW * hp);
        // the marcher waits for the neighbours' flags itself (already satisfied when the
        // exchange was made by k_halo_xchg, which waits)
        t.flag_lo = h.r.flag_lo;
        t.flag_hi = h.r.flag_hi;
        t.err = h.r.err;
        t.dead = h.r.dead;
    } else {
        t.wrap = 0;
        t.halo = c->halo[slot];
        t.halo_fields = 2 * KSFD_SW * hp;
        t.klo = 0;
        t.khi = (int)(KSFD_SW * hp);
    }
    return h;
}

// ---------------------------------------------------------------------------
// Direct halo push over NVLink peer memory.  Every rank owns one IPC-shared
// allocation [64 flag words][slot][parity][lo planes | hi planes]; a single
// kernel writes this rank's top planes into the upper neighbour's lo buffer and
// its bottom planes into the lower neighbour's hi buffer (stores over NVLink),
// publishes the exchange number in the neighbours' flag words and waits for its
// own two flags: no NCCL call, no host involvement, ~one kernel launch of
// latency.  Buffers are double-buffered on the exchange parity: a neighbour may
// already push exchange q+1 while this rank still reads exchange q; every use
// of a slot is separated from its second-next use by a rank-synchronising
// all-reduce (one per Arnoldi step / norm / error norm), so two buffers suffice.
// ---------------------------------------------------------------------------
static size_t p2p_buf_doubles(const ksfd_ctx *c) { return 2 * KSFD_SW * c->halo_plane_doubles; }
static size_t p2p_red_off(const ksfd_ctx *c)
{
    return KSFD_P2P_FLAGS + (size_t)KSFD_HALO_SLOTS * 2 * p2p_buf_doubles(c);
}
static size_t p2p_total_doubles(const ksfd_ctx *c)
{
    // two parities x 16 ranks x 72 doubles, each as two tagged 8-byte words
    return p2p_red_off(c) + (size_t)2 * KSFD_P2P_MAXR * 2 * KSFD_P2P_RED_MAX;
}
static P2PRed p2p_red(const ksfd_ctx *c)
{
    P2PRed pr{};
    pr.nranks = c->p2p_on ? c->nranks : 1;
    pr.rank = c->rank;
    pr.red_off = (long long)p2p_red_off(c);
    pr.ctr = c->p2p_ctr ? c->p2p_ctr + KSFD_HALO_SLOTS : nullptr;
    pr.dead = c->p2p_ctr ? c->p2p_ctr + KSFD_HALO_SLOTS + 1 : nullptr;
    pr.err = c->p2p_err_dev;
    for (int r = 0; r < KSFD_P2P_MAXR; ++r) pr.base[r] = r < c->nranks ? c->p2p_peer[r] : nullptr;
    return pr;
}
static double *p2p_buf(const ksfd_ctx *c, double *base, int slot, int parity)
{
    return base + KSFD_P2P_FLAGS + ((size_t)slot * 2 + parity) * p2p_buf_doubles(c);
}

// up_lo0 / dn_hi0 / flags: parity-0 destinations in the neighbours' allocations;
// ctr = this rank's device-side exchange counter of the slot
__global__ void k_halo_xchg(const double *__restrict__ top, const double *__restrict__ bot,
                            long long cnt, double *__restrict__ up_lo0,
                            double *__restrict__ dn_hi0, long long pstride,
                            volatile unsigned long long *up_flag_lo,
                            volatile unsigned long long *dn_flag_hi,
                            volatile unsigned long long *my_flag_lo,
                            volatile unsigned long long *my_flag_hi, unsigned long long *ctr,
                            unsigned *done, const int *__restrict__ skip, volatile int *err,
                            volatile unsigned long long *dead)
{
    // launched ahead by the pipelined solver: no exchange once the cycle is closed
    if (skip && *skip) return;
    if (*dead) return;                          // an earlier wait timed out: the host reports it
    // every block reads the counter before the last one (which only exists
    // after all blocks have copied) advances it
    const unsigned long long q = *ctr + 1;
    const long long sh = (long long)(q & 1ull) * pstride;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < cnt;
         e += (long long)gridDim.x * blockDim.x) {
        up_lo0[sh + e] = top[e];
        dn_hi0[sh + e] = bot[e];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(done, 1u);
        if (t == gridDim.x - 1) {               // last block: all planes are on their way
            atomicExch(done, 0u);
            __threadfence_system();
            *up_flag_lo = q;
            *dn_flag_hi = q;
            const bool ok = p2p_spin(my_flag_lo, q, err, dead) && p2p_spin(my_flag_hi, q, err, dead);
            __threadfence_system();
            if (ok) *ctr = q;                   // a failed wait does not complete the exchange
        }
    }
}

// push-only variant: the consumer (the TMA-fed marcher, tma_march.cuh: halo_arrived) waits
// for the flags in the CTAs that read ghost planes
__global__ void k_halo_push(const double *__restrict__ top, const double *__restrict__ bot,
                            long long cnt, double *__restrict__ up_lo0,
                            double *__restrict__ dn_hi0, long long pstride,
                            volatile unsigned long long *up_flag_lo,
                            volatile unsigned long long *dn_flag_hi, unsigned long long *ctr,
                            unsigned *done, const int *__restrict__ skip,
                            const volatile unsigned long long *dead)
{
    if (skip && *skip) return;
    if (*dead) return;
    const unsigned long long q = *ctr + 1;
    const long long sh = (long long)(q & 1ull) * pstride;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < cnt;
         e += (long long)gridDim.x * blockDim.x) {
        up_lo0[sh + e] = top[e];
        dn_hi0[sh + e] = bot[e];
    }
    // hierarchical release (halo_push.cuh: halo_push_publish): CTA barrier, one device-scope
    // fence + arrival per block, ONE system-scope fence in the block that arrives last
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned t = atomicAdd(done, 1u);
        if (t == gridDim.x - 1) {
            atomicExch(done, 0u);
            flag_release(up_flag_lo, q);
            flag_release(dn_flag_hi, q);
            *ctr = q;
        }
    }
}


extern "C" int ksfd_p2p_export(ksfd_ctx *c, char handle_out[64])
{
    if (!c || !handle_out) return fail("ksfd_p2p_export: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CK(cudaSetDevice(c->device));
    if (!c->p2p_mine) {
        CK(cudaMalloc(&c->p2p_mine, sizeof(double) * p2p_total_doubles(c)));
        // flags AND the all-reduce area: its words carry their exchange number, 0 = never written
        CK(cudaMemset(c->p2p_mine, 0, sizeof(double) * p2p_total_doubles(c)));
        CK(cudaMalloc(&c->p2p_done, 4 * sizeof(unsigned)));
        CK(cudaMemset(c->p2p_done, 0, 4 * sizeof(unsigned)));
        CK(cudaHostAlloc(&c->p2p_err, sizeof(int), cudaHostAllocMapped));
        *c->p2p_err = 0;
        CK(cudaHostGetDevicePointer(&c->p2p_err_dev, c->p2p_err, 0));
        // [0..3] halo slots, [4] all-reduce, [5] sticky "a peer wait timed out"
        CK(cudaMalloc(&c->p2p_ctr, sizeof(unsigned long long) * (KSFD_HALO_SLOTS + 2)));
        CK(cudaMemset(c->p2p_ctr, 0, sizeof(unsigned long long) * (KSFD_HALO_SLOTS + 2)));
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->p2p_mine));
    memcpy(handle_out, &h, 64);
    return 0;
}

extern "C" int ksfd_p2p_import(ksfd_ctx *c, const char *handles, int nhandles)
{
    if (!c || !handles) return fail("ksfd_p2p_import: NULL argument");
    if (!c->p2p_mine) return fail("ksfd_p2p_import: call ksfd_p2p_export first");
    if (c->nranks < 2) return 0;
    if (nhandles != c->nranks) return fail("ksfd_p2p_import: one handle per rank expected");
    if (c->nranks > KSFD_P2P_MAXR) return fail("ksfd_p2p_import: more than 16 ranks");
    CK(cudaSetDevice(c->device));
    for (int r = 0; r < c->nranks; ++r) {
        if (r == c->rank) {
            c->p2p_peer[r] = c->p2p_mine;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->p2p_peer[r] = static_cast<double *>(p);
    }
    c->p2p_dn = c->p2p_peer[(c->rank + c->nranks - 1) % c->nranks];
    c->p2p_up = c->p2p_peer[(c->rank + 1) % c->nranks];
    c->p2p_on = true;
    return 0;
}

static int exchange_p2p(ksfd_ctx *c, const double *vec, int stride, int slot, cudaStream_t st,
                        const int *skip, bool defer)
{
    const size_t cnt = (size_t)KSFD_SW * c->g.plane_pts * stride;
    const size_t hi_off = KSFD_SW * c->halo_plane_doubles;
    double *up_lo0 = p2p_buf(c, c->p2p_up, slot, 0);               // my top -> up's lo
    double *dn_hi0 = p2p_buf(c, c->p2p_dn, slot, 0) + hi_off;      // my bottom -> dn's hi
    const double *top = vec + (size_t)(c->g.nloc - KSFD_SW) * c->g.plane_pts * stride;
    typedef volatile unsigned long long *flag_t;
    flag_t up_flag_lo = reinterpret_cast<flag_t>(c->p2p_up) + slot * 2 + 0;
    flag_t dn_flag_hi = reinterpret_cast<flag_t>(c->p2p_dn) + slot * 2 + 1;
    flag_t my_flag_lo = reinterpret_cast<flag_t>(c->p2p_mine) + slot * 2 + 0;
    flag_t my_flag_hi = reinterpret_cast<flag_t>(c->p2p_mine) + slot * 2 + 1;
    const unsigned blocks = (unsigned)std::min<size_t>((cnt + 255) / 256, 64);
    if (defer) {
        // the consumer is the TMA-fed marcher: push only, it waits where it reads ghosts
        k_halo_push<<<blocks, 256, 0, st>>>(top, vec, (long long)cnt, up_lo0, dn_hi0,
                                            (long long)p2p_buf_doubles(c), up_flag_lo, dn_flag_hi,
                                            c->p2p_ctr + slot, c->p2p_done, skip,
                                            c->p2p_ctr + KSFD_HALO_SLOTS + 1);
        CKL();
        return 0;
    }
    k_halo_xchg<<<blocks, 256, 0, st>>>(top, vec, (long long)cnt, up_lo0, dn_hi0,
                                        (long long)p2p_buf_doubles(c), up_flag_lo, dn_flag_hi,
                                        my_flag_lo, my_flag_hi, c->p2p_ctr + slot, c->p2p_done,
                                        skip, c->p2p_err_dev, c->p2p_ctr + KSFD_HALO_SLOTS + 1);
    CKL();
    return 0;
}

// Fused producer push (blas1_kernels.cuh: HaloPush) of a dof-strided vector into halo slot
// `slot`; an all-null descriptor when this context does not push (one rank, NCCL
// fallback, or a consumer that is not the TMA-fed marcher)
static bool tma_consumer(const ksfd_ctx *c);
static HaloPush make_push(const ksfd_ctx *c, int slot)
{
    HaloPush hp{};
    if (c->nranks == 1 || !c->p2p_on || !tma_consumer(c) || c->gm_no_push) return hp;
    const long long n = nlocal(c);
    typedef volatile unsigned long long *flag_t;
    hp.up_lo0 = p2p_buf(c, c->p2p_up, slot, 0);
    hp.dn_hi0 = p2p_buf(c, c->p2p_dn, slot, 0) + KSFD_SW * c->halo_plane_doubles;
    hp.pstride = (long long)p2p_buf_doubles(c);
    hp.cnt = (long long)KSFD_SW * c->g.plane_pts * c->dof;
    hp.top0 = n - hp.cnt;
    hp.up_flag_lo = reinterpret_cast<flag_t>(c->p2p_up) + slot * 2 + 0;
    hp.dn_flag_hi = reinterpret_cast<flag_t>(c->p2p_dn) + slot * 2 + 1;
    hp.ctr = c->p2p_ctr + slot;
    hp.done = c->p2p_done + 1;              // own block counter (p2p_done[0]: exchange kernels)
    hp.dead = c->p2p_ctr + KSFD_HALO_SLOTS + 1;
    return hp;
}

static int exchange(ksfd_ctx *c, const double *vec, int stride, int slot,
                    cudaStream_t st, const int *skip = nullptr, bool defer = false)
{
    if (c->nranks == 1) return 0;
    if (slot < 0 || slot >= KSFD_HALO_SLOTS) return fail("bad halo slot");
    // every public entry point that exchanges halos comes through here: a peer wait
    // that timed out earlier (pinned word written by the device) is reported now
    if (c->p2p_err && *c->p2p_err)
        return fail("peer-to-peer exchange timed out (a neighbouring rank is gone)");
    if (c->p2p_on) return exchange_p2p(c, vec, stride, slot, st, skip, defer);
    if (!c->halo[slot])
        CK(cudaMalloc(&c->halo[slot],
                      sizeof(double) * 2 * KSFD_SW * c->halo_plane_doubles));
    const size_t cnt = (size_t)KSFD_SW * c->g.plane_pts * stride;
    const int up = (c->rank + 1) % c->nranks, dn = (c->rank + c->nranks - 1) % c->nranks;
    double *lo = c->halo[slot];
    double *hi = c->halo[slot] + KSFD_SW * c->halo_plane_doubles;
    const double *top = vec + (size_t)(c->g.nloc - KSFD_SW) * c->g.plane_pts * stride;
    // last-axis planes are contiguous (dof fastest, last axis slowest), so a
    // ghost face is one contiguous block: no pack kernel is needed.
    NK(g_nccl.GroupStart());
    NK(g_nccl.Send(top, cnt, ncclFloat64_, up, c->comm, st));     // my top -> up's lo
    NK(g_nccl.Recv(lo, cnt, ncclFloat64_, dn, c->comm, st));
    NK(g_nccl.Send(vec, cnt, ncclFloat64_, dn, c->comm, st));     // my bottom -> dn's hi
    NK(g_nccl.Recv(hi, cnt, ncclFloat64_, up, c->comm, st));
    NK(g_nccl.GroupEnd());
    return 0;
}

extern "C" int ksfd_halo_exchange(ksfd_ctx *c, const double *vec, int slot,
                                  void *stream)
{
    if (!c || !vec) return fail("ksfd_halo_exchange: NULL argument");
    return exchange(c, vec, c->dof, slot, (cudaStream_t)stream);
}

static int allreduce_dev(ksfd_ctx *c, double *buf, int n, int op, cudaStream_t st)
{
    if (c->nranks == 1) return 0;
    if (c->p2p_on && n <= KSFD_P2P_RED_MAX) {
        k_p2p_allreduce<<<1, 128, 0, st>>>(p2p_red(c), buf, n, op == ncclMax_ ? 1 : 0);
        CKL();
        return 0;
    }
    NK(g_nccl.AllReduce(buf, buf, n, ncclFloat64_, op, c->comm, st));
    return 0;
}

// host scalars -> all-reduce on `stream` -> host scalars.  The staging buffers (dscal, the
// peer-to-peer reduce area and its sequence counter) are shared by every reduction of the
// context, so all of them must be issued on ONE stream: the caller's.
static int allreduce_host(ksfd_ctx *c, double *vals, int n, int op, cudaStream_t st,
                          const char *who)
{
    if (!c || !vals || n > 64 || n < 0) return fail(std::string(who) + ": bad argument");
    if (c->nranks == 1 || n == 0) return 0;
    double *stage = c->dscal + KSFD_NSCAL - 64;         // the tail of the scalar scratch
    CK(cudaMemcpyAsync(stage, vals, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    TRY(allreduce_dev(c, stage, n, op, st));
    CK(cudaMemcpyAsync(vals, stage, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (c->p2p_err && *c->p2p_err)
        return fail("peer-to-peer exchange timed out (a neighbouring rank is gone)");
    return 0;
}
extern "C" int ksfd_allreduce_max(ksfd_ctx *c, double *vals, int n, void *stream)
{
    return allreduce_host(c, vals, n, ncclMax_, (cudaStream_t)stream, "ksfd_allreduce_max");
}
extern "C" int ksfd_allreduce_sum(ksfd_ctx *c, double *vals, int n, void *stream)
{
    return allreduce_host(c, vals, n, ncclSum_, (cudaStream_t)stream, "ksfd_allreduce_sum");
}

// ---------------------------------------------------------------------------
// marching kernels: plan cache + dispatch (kernels live in march_*.cu)
// ---------------------------------------------------------------------------

void ksfd_free_plans(ksfd_ctx *c)
{
    delete static_cast<PlanMap *>(c->plan_cache);
    c->plan_cache = nullptr;
    delete static_cast<TmapCache *>(c->tmap_cache);
    c->tmap_cache = nullptr;
}
void ksfd_invalidate_plans(ksfd_ctx *c)
{
    if (c->plan_cache) static_cast<PlanMap *>(c->plan_cache)->clear();
}

static bool use_march(const ksfd_ctx *c);
// both marching kernels wait for the neighbours' flags themselves (the direct kernels do not)
static bool tma_consumer(const ksfd_ctx *c) { return use_march(c); }
static bool use_march(const ksfd_ctx *c)
{
    if (c->variant == 1) return false;
    if (c->dim < 2) return false;
    if (c->dof - 1 > 4) return false;           // instantiated for nlig <= 4
    if (!c->P.sym_ok) return false;             // weights off the symmetric pattern
    if (c->variant == 2) return true;
    return c->g.n0 >= 8 && c->g.nloc >= 4 && (c->dim == 2 || c->g.n1 >= 8);
}

static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

// ---------------------------------------------------------------------------
// operator entry points
// ---------------------------------------------------------------------------
static int check_ready(const ksfd_ctx *c)
{
    if (!c) return fail("NULL context");
    if (!c->have_phys) return fail("ksfd_set_physics has not been called");
    return 0;
}

extern "C" int ksfd_groom(ksfd_ctx *c, double *u, void *stream)
{
    TRY(check_ready(c));
    if (!u) return fail("ksfd_groom: NULL vector");
    k_groom<<<nblk(nlocal(c), 256), 256, 0, (cudaStream_t)stream>>>(
        c->g, c->P.rhomin, c->P.Umin, u);
    CKL();
    return 0;
}

extern "C" int ksfd_to_internal(ksfd_ctx *c, const double *ref, double *out, int nfields,
                                void *stream)
{
    if (!c || !ref || !out || nfields < 1 || ref == out)
        return fail("ksfd_to_internal: bad argument");
    k_to_internal<<<nblk(c->g.npts * nfields, 256), 256, 0, (cudaStream_t)stream>>>(
        c->g, nfields, ref, out);
    CKL();
    return 0;
}

extern "C" int ksfd_from_internal(ksfd_ctx *c, const double *in, double *ref, int nfields,
                                  void *stream)
{
    if (!c || !ref || !in || nfields < 1 || ref == in)
        return fail("ksfd_from_internal: bad argument");
    k_from_internal<<<nblk(c->g.npts * nfields, 256), 256, 0, (cudaStream_t)stream>>>(
        c->g, nfields, in, ref);
    CKL();
    return 0;
}

static HaloPush make_push(const ksfd_ctx *c, int slot);
// push_out: several ranks — the kernel also pushes the boundary planes of f into halo slot 1
// (f is the right-hand side of a solve by Richardson sweeps, whose first sweep reads them)
static int residual_impl(ksfd_ctx *c, const double *u, const double *udot,
                         const double *src, double *f, cudaStream_t st, bool push_out = false)
{
    // ghost planes of u: pushed by the kernel that produced it (k_stage_combine) or here
    if (c->pushed_vec0 == u && u != nullptr)
        c->pushed_vec0 = nullptr;
    else
        TRY(exchange(c, u, c->dof, 0, st, nullptr, c->p2p_on && tma_consumer(c)));
    const HostVec uh = make_hvec(c, u, c->dof, 0);
    ProfScope prof(c, 1, st);
    if (use_march(c)) {
        HaloPush hp{};
        if (push_out && (c->fuse_push_mask & 4)) hp = make_push(c, 1);
        if (hp.up_lo0) c->pushed_vec = f;
        return c->dim == 2 ? ksfd_march_residual_d2(c, uh, udot, src, f, hp.up_lo0 ? &hp : nullptr, st)
                           : ksfd_march_residual_d3(c, uh, udot, src, f, hp.up_lo0 ? &hp : nullptr, st);
    }
    k_residual_naive<<<nblk(c->g.npts, 128), 128, 0, st>>>(c->g, c->P, uh.r, udot,
                                                           src, f);
    CKL();
    return 0;
}

extern "C" int ksfd_residual(ksfd_ctx *c, const double *u, const double *udot,
                             const double *src, double *f, void *stream)
{
    TRY(check_ready(c));
    if (!u || !f) return fail("ksfd_residual: NULL vector");
    return residual_impl(c, u, udot, src, f, (cudaStream_t)stream);
}

static int velocity_impl(ksfd_ctx *c, const double *u, double *vel, double *vmax,
                         cudaStream_t st)
{
    TRY(exchange(c, u, c->dof, 0, st, nullptr, c->p2p_on && tma_consumer(c)));
    const HostVec uh = make_hvec(c, u, c->dof, 0);
    if (vmax) CK(cudaMemsetAsync(vmax, 0, sizeof(double) * c->dim, st));
    if (use_march(c)) {
        return c->dim == 2 ? ksfd_march_velocity_d2(c, uh, vel, vmax, st)
                           : ksfd_march_velocity_d3(c, uh, vel, vmax, st);
    }
    k_velocity_naive<<<nblk(c->g.npts, 128), 128, 0, st>>>(c->g, c->P, uh.r, vel, vmax);
    CKL();
    return 0;
}

extern "C" int ksfd_velocity_max(ksfd_ctx *c, const double *u, double *vmax,
                                 void *stream)
{
    TRY(check_ready(c));
    if (!u || !vmax) return fail("ksfd_velocity_max: NULL argument");
    return velocity_impl(c, u, nullptr, vmax, (cudaStream_t)stream);
}

extern "C" int ksfd_velocity(ksfd_ctx *c, const double *u, double *vel, void *stream)
{
    TRY(check_ready(c));
    if (!u || !vel) return fail("ksfd_velocity: NULL argument");
    return velocity_impl(c, u, vel, nullptr, (cudaStream_t)stream);
}

// coef is stored ghosted along the last axis: base = coef + 2 planes
static VecRef coef_ref(const ksfd_ctx *c)
{
    const long long ps = c->g.plane_pts * (c->dof + 2);
    VecRef r;
    r.par = nullptr;
    r.pstride = 0;
    r.lo = c->coef;
    r.base = c->coef + KSFD_SW * ps;
    r.hi = c->coef + (long long)(KSFD_SW + c->g.nloc) * ps;
    return r;
}

// the coefficient field for the TMA-fed marcher: one ghosted buffer, ghost planes in place
static HostVec coef_hvec(const ksfd_ctx *c)
{
    HostVec h;
    h.r = coef_ref(c);
    const int nf = c->dof + 2;
    h.t.base = c->coef;
    h.t.base_fields = (long long)(c->g.nloc + 2 * KSFD_SW) * nf;
    h.t.k0 = KSFD_SW * nf;
    h.t.halo = nullptr;
    h.t.wrap = 0;
    h.t.klo = 0;
    h.t.khi = (KSFD_SW + c->g.nloc) * nf;
    return h;
}

// ghosts_current: the ghost planes of u in halo slot 0 are those of the exchange the stage
// residual just consumed (same vector, nothing in between): no second exchange
static int jvp_setup_impl(ksfd_ctx *c, const double *u, double shift,
                          double *blocks, cudaStream_t st, bool ghosts_current = false)
{
    const Geom &g = c->g;
    const long long gpts = (long long)(g.nloc + 2 * KSFD_SW) * g.plane_pts;
    if (!c->coef) CK(cudaMalloc(&c->coef, sizeof(double) * gpts * (c->dof + 2)));
    if (!c->pc) CK(cudaMalloc(&c->pc, sizeof(double) * g.npts));
    c->shift = shift;
    c->Pjac = c->P;
    for (int l = 0; l < c->P.nlig; ++l)
        c->invd[l] = 1.0 / (shift + c->P.gamma[l] - c->P.D[l] * c->P.w2c);
    if (u) {
        if (!(ghosts_current && c->p2p_on)) TRY(exchange(c, u, c->dof, 0, st));
        VecRef ur = make_ref(c, u, c->dof, 0);
        k_coef_setup<<<nblk(gpts, 128), 128, 0, st>>>(g, c->P, ur, c->coef);
        CKL();
    }
    InvD id;
    for (int l = 0; l < KSFD_MAX_LIGANDS; ++l) id.v[l] = c->invd[l];
    k_pc_setup<<<nblk(g.npts, 128), 128, 0, st>>>(g, c->P, coef_ref(c), shift, id, c->pc,
                                                  blocks);
    CKL();
    // ghost planes of the preconditioner field for the fused A*M^{-1} kernel: push only where
    // the consumers are the marching kernels (they wait for the neighbours' flags themselves)
    TRY(exchange(c, c->pc, 1, 2, st, nullptr, c->p2p_on && tma_consumer(c)));
    c->have_jac = true;
    c->fft_means_valid = false;
    return 0;
}

extern "C" int ksfd_jvp_setup(ksfd_ctx *c, const double *u, double shift, void *stream)
{
    TRY(check_ready(c));
    if (!u) return fail("ksfd_jvp_setup: NULL vector");
    return jvp_setup_impl(c, u, shift, nullptr, (cudaStream_t)stream);
}

extern "C" int ksfd_block_diagonal(ksfd_ctx *c, double *blocks, void *stream)
{
    TRY(check_ready(c));
    if (!c->have_jac) return fail("ksfd_jvp_setup has not been called");
    CK(cudaMemsetAsync(blocks, 0, sizeof(double) * c->g.npts * c->dof * c->dof,
                       (cudaStream_t)stream));
    return jvp_setup_impl(c, nullptr, c->shift, blocks, (cudaStream_t)stream);
}

static int jvp_impl(ksfd_ctx *c, const double *v, double *out, bool precond,
                    cudaStream_t st, const int *skip = nullptr)
{
    if (!c->have_jac) return fail("ksfd_jvp_setup has not been called");
    if (v == out) return fail("ksfd_jvp: in-place application is not supported");
    // several ranks over peer memory + TMA-fed marcher: nobody waits for the halo on the
    // stream.  Either the producer of v already pushed its boundary planes (Krylov
    // vectors: fused into k_gm_first_vector / k_gm_orth_scale, c->pushed_vec), or a
    // push-only kernel goes out; the marcher waits in the CTAs that read ghost planes.
    const bool defer = c->p2p_on && c->nranks > 1 && tma_consumer(c);
    if (defer && v == c->pushed_vec) {
        c->pushed_vec = nullptr;            // consumed: pushed by its producer
    } else {
        TRY(exchange(c, v, c->dof, 1, st, skip, defer));
    }
    HostVec vh = make_hvec(c, v, c->dof, 1);
    const HostVec ph = make_hvec(c, c->pc, 1, 2);
    const HostVec ch = coef_hvec(c);
    VecRef &vr = vh.r;
    const VecRef &pr = ph.r, &cr = ch.r;
    ProfScope prof(c, 0, st);
    if (use_march(c)) {
        return c->dim == 2 ? ksfd_march_jvp_d2(c, ch, vh, ph, precond, out, skip, st)
                           : ksfd_march_jvp_d3(c, ch, vh, ph, precond, out, skip, st);
    }
    InvD id;
    for (int l = 0; l < KSFD_MAX_LIGANDS; ++l) id.v[l] = c->invd[l];
    k_jvp_naive<<<nblk(c->g.npts, 128), 128, 0, st>>>(c->g, c->Pjac, cr, vr, pr, id,
                                                      precond ? 1 : 0, c->shift, skip, out);
    CKL();
    return 0;
}

extern "C" int ksfd_jvp(ksfd_ctx *c, const double *v, double *out, void *stream)
{
    TRY(check_ready(c));
    if (!v || !out) return fail("ksfd_jvp: NULL vector");
    return jvp_impl(c, v, out, false, (cudaStream_t)stream);
}
extern "C" int ksfd_jvp_precond(ksfd_ctx *c, const double *v, double *out, void *stream)
{
    TRY(check_ready(c));
    if (!v || !out) return fail("ksfd_jvp_precond: NULL vector");
    return jvp_impl(c, v, out, true, (cudaStream_t)stream);
}

static int pc_apply_impl(ksfd_ctx *c, const double *r, double *z, cudaStream_t st)
{
    InvD id;
    for (int l = 0; l < KSFD_MAX_LIGANDS; ++l) id.v[l] = c->invd[l];
    k_pc_apply<<<nblk(c->g.npts, 128), 128, 0, st>>>(c->g, c->Pjac, coef_ref(c), id, c->pc, r, z);
    CKL();
    return 0;
}

extern "C" int ksfd_pc_apply(ksfd_ctx *c, const double *r, double *z, void *stream)
{
    TRY(check_ready(c));
    if (!c->have_jac) return fail("ksfd_jvp_setup has not been called");
    if (!r || !z) return fail("ksfd_pc_apply: NULL vector");
    return pc_apply_impl(c, r, z, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
// spectral preconditioner (fftpc.cuh): cuFFT through dlopen (CUDA toolkit's
// libcufft.so.11; no link-time dependency), plans created once per context
// ---------------------------------------------------------------------------
struct CufftApi {
    void *h = nullptr;
    bool tried = false;
    int (*PlanMany)(int *, int, int *, int *, int, int, int *, int, int, int, int) = nullptr;
    int (*SetStream)(int, cudaStream_t) = nullptr;
    int (*ExecD2Z)(int, double *, double2 *) = nullptr;
    int (*ExecZ2D)(int, double2 *, double *) = nullptr;
    int (*ExecZ2Z)(int, double2 *, double2 *, int) = nullptr;
    int (*Destroy)(int) = nullptr;
};
static CufftApi g_fft;
enum { CUFFT_D2Z_ = 0x6a, CUFFT_Z2D_ = 0x6c, CUFFT_Z2Z_ = 0x69 };

static bool cufft_load()
{
    if (g_fft.tried) return g_fft.h != nullptr;
    g_fft.tried = true;
    const char *names[] = {"libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so.11", "libcufft.so",
                           "libcufft.so.12"};
    void *h = nullptr;
    for (const char *nm : names)
        if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return false;
#define FSYM(n)                                        \
    *(void **)(&g_fft.n) = dlsym(h, "cufft" #n);       \
    if (!g_fft.n) return false;
    FSYM(PlanMany) FSYM(SetStream) FSYM(ExecD2Z) FSYM(ExecZ2D) FSYM(ExecZ2Z) FSYM(Destroy)
#undef FSYM
    g_fft.h = h;
    return true;
}

static void fftpc_destroy(ksfd_ctx *c)
{
    if (c->fft_fwd >= 0 && g_fft.Destroy) {
        g_fft.Destroy(c->fft_fwd);
        g_fft.Destroy(c->fft_inv);
        if (c->fft_z >= 0) g_fft.Destroy(c->fft_z);
        c->fft_fwd = c->fft_inv = c->fft_z = -1;
    }
}

// Several ranks: slab-distributed transform (plane transforms of the own planes, one
// all-to-all per direction over NCCL send/recv, last-axis transform; fftpc.cuh), 2-D / 3-D,
// with the DMDA ownership ranges (every rank must be able to compute every other rank's).
// Validated on 2 B200 against the single-GPU preconditioner: same Arnoldi counts, solutions
// to 1e-16 (scripts/multi_gpu_spectral_check.py, tests/test_gpu_multi.py);
// KSFD_FFT_MULTI=0 switches it off (block Jacobi is used instead).
static bool fftpc_dist_wanted(const ksfd_ctx *c)
{
    if (c->nranks == 1) return false;
    const char *e = getenv("KSFD_FFT_MULTI");
    if ((e && atoi(e) == 0) || !c->comm || c->dim < 2) return false;
    const long long M = c->last_global, P = c->nranks, r = c->rank;
    const long long cnt = M / P + (M % P > r ? 1 : 0);
    const long long start = r * (M / P) + std::min(r, M % P);
    return cnt == c->last_count && start == c->last_start;
}

static bool fftpc_available(ksfd_ctx *c)
{
    return (c->nranks == 1 || fftpc_dist_wanted(c)) && !c->fft_failed && cufft_load();
}

// plans + buffers (first use); returns false (and remembers) if cuFFT refuses
static bool fftpc_prepare(ksfd_ctx *c)
{
    if (c->fft_fwd >= 0) return true;
    if (!fftpc_available(c)) return false;
    const int dof = c->dof;
    const int nx = (int)c->n[0], ny = c->dim >= 2 ? (int)c->n[1] : 1, nz = c->dim >= 3 ? (int)c->n[2] : 1;
    const int nxh = nx / 2 + 1;
    if (c->nranks > 1) {
        // plane transforms over the own planes, 1-D transform along the last axis
        const int nloc = c->g.nloc, NL = (int)c->last_global, P = c->nranks;
        const long long PS = c->dim == 2 ? nxh : (long long)ny * nxh;
        const long long nsq = fft_share_start(PS, c->rank + 1, P) - fft_share_start(PS, c->rank, P);
        int f = -1, b = -1, z = -1;
        int np[2], ie[2], oe[2], prank;
        if (c->dim == 2) {
            prank = 1; np[0] = nx; ie[0] = nx; oe[0] = nxh;
        } else {
            prank = 2; np[0] = ny; np[1] = nx; ie[0] = ny; ie[1] = nx; oe[0] = ny; oe[1] = nxh;
        }
        const int pp = (int)c->g.plane_pts;
        int nz1[1] = {NL}, ez[1] = {NL};
        bool ok = nsq > 0 &&
                  g_fft.PlanMany(&f, prank, np, ie, 1, pp, oe, 1, (int)PS, CUFFT_D2Z_, nloc * dof) == 0 &&
                  g_fft.PlanMany(&b, prank, np, oe, 1, (int)PS, ie, 1, pp, CUFFT_Z2D_, nloc * dof) == 0 &&
                  g_fft.PlanMany(&z, 1, nz1, ez, 1, NL, ez, 1, NL, CUFFT_Z2Z_, (int)(dof * nsq)) == 0;
        const size_t cap = (size_t)std::max((long long)nloc * dof * PS, (long long)dof * nsq * NL);
        ok = ok && cudaMalloc(&c->fft_spec, sizeof(double2) * cap) == cudaSuccess &&
             cudaMalloc(&c->fft_spec2, sizeof(double2) * cap) == cudaSuccess &&
             cudaMalloc(&c->fft_means,
                        sizeof(double) * (KSFD_MAX_LIGANDS + 2) * (FFT_MEAN_BLOCKS + 1)) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            if (f >= 0) g_fft.Destroy(f);
            if (b >= 0) g_fft.Destroy(b);
            if (z >= 0) g_fft.Destroy(z);
            c->fft_failed = true;
            return false;
        }
        c->fft_fwd = f;
        c->fft_inv = b;
        c->fft_z = z;
        c->fft_dist = true;
        c->fft_ps = PS;
        return true;
    }
    int rank = c->dim, n[3], ie[3], oe[3], istride = 1, idist, odist;
    if (c->dim == 1) {
        n[0] = nx; ie[0] = nx; oe[0] = nxh; istride = dof; idist = 1; odist = nxh;
    } else if (c->dim == 2) {
        n[0] = ny; n[1] = nx; ie[0] = ny; ie[1] = dof * nx; oe[0] = ny; oe[1] = nxh;
        idist = nx; odist = ny * nxh;
    } else {
        n[0] = nz; n[1] = ny; n[2] = nx; ie[0] = nz; ie[1] = dof * ny; ie[2] = nx;
        oe[0] = nz; oe[1] = ny; oe[2] = nxh; idist = ny * nx; odist = nz * ny * nxh;
    }
    int f = -1, b = -1;
    if (g_fft.PlanMany(&f, rank, n, ie, istride, idist, oe, 1, odist, CUFFT_D2Z_, dof) != 0 ||
        g_fft.PlanMany(&b, rank, n, oe, 1, odist, ie, istride, idist, CUFFT_Z2D_, dof) != 0) {
        if (f >= 0) g_fft.Destroy(f);
        c->fft_failed = true;
        return false;
    }
    const size_t nspec = (size_t)dof * nz * ny * nxh;
    if (cudaMalloc(&c->fft_spec, sizeof(double2) * nspec) != cudaSuccess ||
        cudaMalloc(&c->fft_means, sizeof(double) * (KSFD_MAX_LIGANDS + 2) * (FFT_MEAN_BLOCKS + 1)) !=
            cudaSuccess) {
        cudaGetLastError();
        g_fft.Destroy(f);
        g_fft.Destroy(b);
        c->fft_failed = true;
        return false;
    }
    c->fft_fwd = f;
    c->fft_inv = b;
    return true;
}

static VecRef coef_ref(const ksfd_ctx *c);

// coefficient means of the current linearisation (after k_coef_setup)
static int fftpc_setup(ksfd_ctx *c, cudaStream_t st)
{
    if (!fftpc_prepare(c)) return 0;
    double *partial = c->fft_means + (KSFD_MAX_LIGANDS + 2);
    k_fft_means_partial<<<dim3(c->P.nlig + 2, FFT_MEAN_BLOCKS), 256, 0, st>>>(c->g, coef_ref(c).base,
                                                                           partial);
    CKL();
    double gcount = 1.0;
    for (int a = 0; a < c->dim; ++a) gcount *= (double)c->n[a];
    k_fft_means_final<<<c->P.nlig + 2, 32, 0, st>>>(1.0 / gcount, partial, c->fft_means);
    CKL();
    if (c->nranks > 1) TRY(allreduce_dev(c, c->fft_means, c->P.nlig + 2, ncclSum_, st));
    return 0;
}

// out = S_R A0^-1 S_L in   (in, out: distinct plane-SoA vectors; scratch holds the
// row-scaled copy of `in` and may be `in` itself if that may be overwritten)
static int fftpc_apply(ksfd_ctx *c, const double *in, double *scratch, double *out,
                       cudaStream_t st, const int *skip)
{
    FftSym S{};
    S.dof = c->dof;
    S.nlig = c->Pjac.nlig;
    S.n0 = (int)c->n[0];
    S.n1 = c->dim >= 2 ? (int)c->n[1] : 1;
    S.n2 = c->dim >= 3 ? (int)c->n[2] : 1;
    S.shift = c->shift;
    for (int a = 0; a < 3; ++a) S.c2[a] = c->Pjac.c2[a];
    for (int l = 0; l < KSFD_MAX_LIGANDS; ++l) {
        S.s[l] = c->Pjac.s[l];
        S.gamma[l] = c->Pjac.gamma[l];
        S.D[l] = c->Pjac.D[l];
    }
    double2 *spec = static_cast<double2 *>(c->fft_spec);
    if (g_fft.SetStream(c->fft_fwd, st) != 0 || g_fft.SetStream(c->fft_inv, st) != 0)
        return fail("cufftSetStream failed");
    k_fft_prescale<<<nblk(nlocal(c), 256), 256, 0, st>>>(c->g, coef_ref(c).base, in, scratch, skip);
    CKL();
    if (c->fft_dist) {
        // slab-distributed transform (fftpc.cuh): A = plane spectra / receive buffer,
        // B = packed send buffer / last-axis-contiguous spectra.  Every rank makes the
        // same NCCL calls whether or not the kernels skip.
        double2 *A = spec, *B = static_cast<double2 *>(c->fft_spec2);
        const int P = c->nranks, nloc = c->g.nloc, dof = c->dof, NL = (int)c->last_global;
        const long long PS = c->fft_ps;
        const long long s0 = fft_share_start(PS, c->rank, P);
        const long long nsq = fft_share_start(PS, c->rank + 1, P) - s0;
        auto k0_of = [&](int r) { return (long long)r * (NL / P) + std::min<long long>(r, NL % P); };
        auto all_to_all = [&](double2 *send, double2 *recv, bool fwd) -> int {
            // fwd: my planes' share q -> rank q;  back: rank p's planes of my share -> rank p
            NK(g_nccl.GroupStart());
            for (int r = 0; r < P; ++r) {
                const long long sr = fft_share_start(PS, r, P);
                const long long nsr = fft_share_start(PS, r + 1, P) - sr;
                const long long kr = k0_of(r), nlr = k0_of(r + 1) - kr;
                const long long by_share = (long long)nloc * dof * sr, n_share = (long long)nloc * dof * nsr;
                const long long by_plane = (long long)dof * nsq * kr, n_plane = nlr * dof * nsq;
                if (fwd) {
                    NK(g_nccl.Send(send + by_share, 2 * n_share, ncclFloat64_, r, c->comm, st));
                    NK(g_nccl.Recv(recv + by_plane, 2 * n_plane, ncclFloat64_, r, c->comm, st));
                } else {
                    NK(g_nccl.Send(send + by_plane, 2 * n_plane, ncclFloat64_, r, c->comm, st));
                    NK(g_nccl.Recv(recv + by_share, 2 * n_share, ncclFloat64_, r, c->comm, st));
                }
            }
            NK(g_nccl.GroupEnd());
            return 0;
        };
        if (g_fft.SetStream(c->fft_z, st) != 0) return fail("cufftSetStream failed");
        if (g_fft.ExecD2Z(c->fft_fwd, scratch, A) != 0) return fail("cufftExecD2Z failed");
        const long long nA = (long long)nloc * dof * PS, nT = (long long)NL * dof * nsq;
        k_fft_pack<<<nblk(nA, 256), 256, 0, st>>>(nloc, dof, PS, P, 1, A, B, skip);
        CKL();
        TRY(all_to_all(B, A, true));
        k_fft_transpose<<<nblk(nT, 256), 256, 0, st>>>(NL, dof, nsq, 1, A, B, skip);
        CKL();
        if (g_fft.ExecZ2Z(c->fft_z, B, B, -1) != 0) return fail("cufftExecZ2Z failed");
        S.dist = 1;
        S.s0 = (int)s0;
        S.nsq = (int)nsq;
        S.NL = NL;
        k_fft_symbol_solve<<<nblk(nsq * NL, 256), 256, 0, st>>>(S, c->fft_means, B, skip);
        CKL();
        if (g_fft.ExecZ2Z(c->fft_z, B, B, 1) != 0) return fail("cufftExecZ2Z failed");
        k_fft_transpose<<<nblk(nT, 256), 256, 0, st>>>(NL, dof, nsq, 0, A, B, skip);
        CKL();
        TRY(all_to_all(A, B, false));
        k_fft_pack<<<nblk(nA, 256), 256, 0, st>>>(nloc, dof, PS, P, 0, A, B, skip);
        CKL();
        if (g_fft.ExecZ2D(c->fft_inv, A, out) != 0) return fail("cufftExecZ2D failed");
        k_fft_postscale<<<nblk(c->g.npts, 256), 256, 0, st>>>(c->g, coef_ref(c).base, out, skip);
        CKL();
        return 0;
    }
    if (g_fft.ExecD2Z(c->fft_fwd, scratch, spec) != 0) return fail("cufftExecD2Z failed");
    const long long nk = (long long)(S.n0 / 2 + 1) * S.n1 * S.n2;
    k_fft_symbol_solve<<<nblk(nk, 256), 256, 0, st>>>(S, c->fft_means, spec, skip);
    CKL();
    if (g_fft.ExecZ2D(c->fft_inv, spec, out) != 0) return fail("cufftExecZ2D failed");
    k_fft_postscale<<<nblk(c->g.npts, 256), 256, 0, st>>>(c->g, coef_ref(c).base, out, skip);
    CKL();
    return 0;
}

// ---------------------------------------------------------------------------
// BLAS-1
// ---------------------------------------------------------------------------
template <int NV>
static int mdot_launch(ksfd_ctx *c, const VecList &vl, const double *w, cudaStream_t st)
{
    k_mdot<NV><<<KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st>>>(nlocal(c), vl, w, c->partial);
    CKL();
    return 0;
}

// out_dev[i] = <vs[i], w>, i < nv (any nv); global over ranks
static int mdot_impl(ksfd_ctx *c, int nv, const double *const *vs, const double *w,
                     double *out_dev, cudaStream_t st)
{
    for (int b = 0; b < nv; b += KSFD_MAXV) {
        const int m = std::min(KSFD_MAXV, nv - b);
        VecList vl;
        for (int i = 0; i < KSFD_MAXV; ++i) vl.v[i] = vs[b + std::min(i, m - 1)];
        switch (m) {
        case 1: TRY(mdot_launch<1>(c, vl, w, st)); break;
        case 2: TRY(mdot_launch<2>(c, vl, w, st)); break;
        case 3: TRY(mdot_launch<3>(c, vl, w, st)); break;
        case 4: TRY(mdot_launch<4>(c, vl, w, st)); break;
        case 5: TRY(mdot_launch<5>(c, vl, w, st)); break;
        case 6: TRY(mdot_launch<6>(c, vl, w, st)); break;
        case 7: TRY(mdot_launch<7>(c, vl, w, st)); break;
        default: TRY(mdot_launch<8>(c, vl, w, st)); break;
        }
        k_reduce_partials<<<m, 128, 0, st>>>(m, KSFD_RED_BLOCKS, c->partial, out_dev + b, 0);
        CKL();
    }
    TRY(allreduce_dev(c, out_dev, nv, ncclSum_, st));
    return 0;
}

extern "C" int ksfd_mdot(ksfd_ctx *c, int nv, const double *const *vs,
                         const double *w, double *out_dev, void *stream)
{
    if (!c || nv < 1 || !vs || !w || !out_dev) return fail("ksfd_mdot: bad argument");
    return mdot_impl(c, nv, vs, w, out_dev, (cudaStream_t)stream);
}

template <int NV>
static int maxpy_launch(ksfd_ctx *c, const CoefList &cl, const VecList &vl, double ys,
                        double *y, cudaStream_t st)
{
    k_maxpy_host<NV><<<KSFD_RED_BLOCKS, 256, 0, st>>>(nlocal(c), cl, vl, ys, y);
    CKL();
    return 0;
}

// y = yscale*y + sum coef[i]*vs[i]
static int maxpy_impl(ksfd_ctx *c, int nv, const double *coef, const double *const *vs,
                      double yscale, double *y, cudaStream_t st)
{
    if (nv == 0) {
        if (yscale == 1.0) return 0;
        CoefList cl{};
        VecList vl{};
        vl.v[0] = y;
        cl.c[0] = 0.0;
        return maxpy_launch<1>(c, cl, vl, yscale, y, st);
    }
    for (int b = 0; b < nv; b += KSFD_MAXV) {
        const int m = std::min(KSFD_MAXV, nv - b);
        CoefList cl{};
        VecList vl{};
        for (int i = 0; i < m; ++i) {
            cl.c[i] = coef[b + i];
            vl.v[i] = vs[b + i];
        }
        const double ys = (b == 0) ? yscale : 1.0;
        switch (m) {
        case 1: TRY(maxpy_launch<1>(c, cl, vl, ys, y, st)); break;
        case 2: TRY(maxpy_launch<2>(c, cl, vl, ys, y, st)); break;
        case 3: TRY(maxpy_launch<3>(c, cl, vl, ys, y, st)); break;
        case 4: TRY(maxpy_launch<4>(c, cl, vl, ys, y, st)); break;
        case 5: TRY(maxpy_launch<5>(c, cl, vl, ys, y, st)); break;
        case 6: TRY(maxpy_launch<6>(c, cl, vl, ys, y, st)); break;
        case 7: TRY(maxpy_launch<7>(c, cl, vl, ys, y, st)); break;
        default: TRY(maxpy_launch<8>(c, cl, vl, ys, y, st)); break;
        }
    }
    return 0;
}

extern "C" int ksfd_maxpy(ksfd_ctx *c, int nv, const double *coef,
                          const double *const *vs, double *y, void *stream)
{
    if (!c || nv < 0 || !y) return fail("ksfd_maxpy: bad argument");
    return maxpy_impl(c, nv, coef, vs, 1.0, y, (cudaStream_t)stream);
}

// global 2-norm -> dscal[slot] on the device
static int norm2_dev(ksfd_ctx *c, const double *x, int slot, cudaStream_t st)
{
    VecList vl;
    for (int i = 0; i < KSFD_MAXV; ++i) vl.v[i] = x;
    k_mdot<1><<<KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st>>>(nlocal(c), vl, x, c->partial);
    CKL();
    if (c->nranks == 1) {
        k_reduce_partials<<<1, 128, 0, st>>>(1, KSFD_RED_BLOCKS, c->partial, c->dscal + slot, 2);
        CKL();
    } else {
        k_reduce_partials<<<1, 128, 0, st>>>(1, KSFD_RED_BLOCKS, c->partial, c->dscal + slot, 0);
        CKL();
        TRY(allreduce_dev(c, c->dscal + slot, 1, ncclSum_, st));
        k_reduce_partials<<<1, 32, 0, st>>>(1, 1, c->dscal + slot, c->dscal + slot, 2);
        CKL();
    }
    return 0;
}

static int fetch(ksfd_ctx *c, int slot, int n, cudaStream_t st)
{
    CK(cudaMemcpyAsync(c->hscal + slot, c->dscal + slot, sizeof(double) * n,
                       cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (c->p2p_err && *c->p2p_err)
        return fail("peer-to-peer exchange timed out (a neighbouring rank is gone)");
    return 0;
}

extern "C" int ksfd_norm2(ksfd_ctx *c, const double *x, double *out, void *stream)
{
    if (!c || !x || !out) return fail("ksfd_norm2: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    TRY(norm2_dev(c, x, SC_USER, st));
    TRY(fetch(c, SC_USER, 1, st));
    *out = c->hscal[SC_USER];
    return 0;
}

extern "C" int ksfd_sum_dof0(ksfd_ctx *c, const double *u, double *out, void *stream)
{
    if (!c || !u || !out) return fail("ksfd_sum_dof0: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    k_sum_dof0<<<KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st>>>(c->g.npts, c->g.plane_pts, c->dof, u,
                                                             c->partial);
    CKL();
    k_reduce_partials<<<1, 128, 0, st>>>(1, KSFD_RED_BLOCKS, c->partial, c->dscal + SC_USER + 1, 0);
    CKL();
    TRY(allreduce_dev(c, c->dscal + SC_USER + 1, 1, ncclSum_, st));
    TRY(fetch(c, SC_USER + 1, 1, st));
    *out = c->hscal[SC_USER + 1];
    return 0;
}

extern "C" int ksfd_mul_exp_dof0(ksfd_ctx *c, double *u, const double *z, double sd, void *stream)
{
    if (!c || !u || !z) return fail("ksfd_mul_exp_dof0: bad argument");
    k_mul_exp_dof0<<<KSFD_RED_BLOCKS, 256, 0, (cudaStream_t)stream>>>(c->g.npts, c->g.plane_pts,
                                                                      c->dof, sd, z, u);
    CKL();
    return 0;
}

extern "C" int ksfd_scale_dof0(ksfd_ctx *c, double *u, double f, void *stream)
{
    if (!c || !u) return fail("ksfd_scale_dof0: bad argument");
    k_scale_dof0<<<KSFD_RED_BLOCKS, 256, 0, (cudaStream_t)stream>>>(c->g.npts, c->g.plane_pts, c->dof, f, u);
    CKL();
    return 0;
}

// ---------------------------------------------------------------------------
// GMRES(m), right preconditioned by point-block Jacobi, device resident:
// one fused A*M^{-1} stencil kernel, one fused multi-dot, one fused
// orthogonalise+norm kernel per iteration; the host only sees the (j+2)
// Hessenberg entries of the new column (one small D2H copy per iteration).
// ---------------------------------------------------------------------------
template <int NV>
static int orth_launch(ksfd_ctx *c, const VecList &vl, const double *h, double *w,
                       int want_norm, cudaStream_t st)
{
    k_orth_update<NV><<<KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st>>>(nlocal(c), vl, h, w,
                                                                   c->partial, want_norm);
    CKL();
    return 0;
}

static int orth_impl(ksfd_ctx *c, int nv, double *V, const double *h_dev, double *w,
                     double *norm_dev, cudaStream_t st)
{
    const long long n = nlocal(c);
    for (int b = 0; b < nv; b += KSFD_MAXV) {
        const int m = std::min(KSFD_MAXV, nv - b);
        const int last = (b + m >= nv);
        VecList vl;
        for (int i = 0; i < KSFD_MAXV; ++i) vl.v[i] = V + (long long)(b + std::min(i, m - 1)) * n;
        switch (m) {
        case 1: TRY(orth_launch<1>(c, vl, h_dev + b, w, last, st)); break;
        case 2: TRY(orth_launch<2>(c, vl, h_dev + b, w, last, st)); break;
        case 3: TRY(orth_launch<3>(c, vl, h_dev + b, w, last, st)); break;
        case 4: TRY(orth_launch<4>(c, vl, h_dev + b, w, last, st)); break;
        case 5: TRY(orth_launch<5>(c, vl, h_dev + b, w, last, st)); break;
        case 6: TRY(orth_launch<6>(c, vl, h_dev + b, w, last, st)); break;
        case 7: TRY(orth_launch<7>(c, vl, h_dev + b, w, last, st)); break;
        default: TRY(orth_launch<8>(c, vl, h_dev + b, w, last, st)); break;
        }
    }
    if (c->nranks == 1) {
        k_reduce_partials<<<1, 128, 0, st>>>(1, KSFD_RED_BLOCKS, c->partial, norm_dev, 2);
        CKL();
    } else {
        k_reduce_partials<<<1, 128, 0, st>>>(1, KSFD_RED_BLOCKS, c->partial, norm_dev, 0);
        CKL();
        TRY(allreduce_dev(c, norm_dev, 1, ncclSum_, st));
        k_reduce_partials<<<1, 32, 0, st>>>(1, 1, norm_dev, norm_dev, 2);
        CKL();
    }
    return 0;
}

template <int NV>
static int orth_scale_launch(ksfd_ctx *c, const VecList &vl, const double *h,
                             const double *inv, double *w, cudaStream_t st)
{
    k_orth_scale<NV><<<KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st>>>(nlocal(c), vl, h, inv, w);
    CKL();
    return 0;
}

// One classical Gram-Schmidt Arnoldi step with the norm of the new vector
// obtained from <w,w> - sum h_i^2 (no second pass over w):
//   mdot over (V_0..V_j, w) -> h[0..j], <w,w> ; finalize -> h[j+1], 1/h[j+1] ;
//   w <- (w - V h)/h[j+1]   in one pass.
// Leaves h[0..j+1] in dscal[SC_H..] and [inv, flag] in dscal[SC_AUX..].
static int cgs_fused_step(ksfd_ctx *c, int j, double *V, double *w, cudaStream_t st)
{
    const long long n = nlocal(c);
    double *hdev = c->dscal + SC_H;
    double *aux = c->dscal + SC_AUX;
    const int nv = j + 2;                     // V_0..V_j and w itself
    const double thresh = 1e-4;
    for (int b = 0; b < nv; b += KSFD_MAXV) {
        const int m = std::min(KSFD_MAXV, nv - b);
        const bool last = b + m >= nv;
        VecList vl;
        for (int i = 0; i < KSFD_MAXV; ++i) {
            const int idx = b + std::min(i, m - 1);
            vl.v[i] = (idx == j + 1) ? w : V + (long long)idx * n;
        }
        switch (m) {
        case 1: TRY(mdot_launch<1>(c, vl, w, st)); break;
        case 2: TRY(mdot_launch<2>(c, vl, w, st)); break;
        case 3: TRY(mdot_launch<3>(c, vl, w, st)); break;
        case 4: TRY(mdot_launch<4>(c, vl, w, st)); break;
        case 5: TRY(mdot_launch<5>(c, vl, w, st)); break;
        case 6: TRY(mdot_launch<6>(c, vl, w, st)); break;
        case 7: TRY(mdot_launch<7>(c, vl, w, st)); break;
        default: TRY(mdot_launch<8>(c, vl, w, st)); break;
        }
        if (last && c->nranks == 1) {
            k_gs_finalize<<<1, 256, 0, st>>>(m, b, j + 1, KSFD_RED_BLOCKS, c->partial, hdev,
                                             aux, thresh);
            CKL();
        } else {
            k_reduce_partials<<<m, 128, 0, st>>>(m, KSFD_RED_BLOCKS, c->partial, hdev + b, 0);
            CKL();
        }
    }
    if (c->nranks > 1) {
        TRY(allreduce_dev(c, hdev, nv, ncclSum_, st));
        k_gs_finalize_only<<<1, 32, 0, st>>>(j + 1, hdev, aux, thresh);
        CKL();
    }
    const int no = j + 1;
    for (int b = 0; b < no; b += KSFD_MAXV) {
        const int m = std::min(KSFD_MAXV, no - b);
        const bool last = b + m >= no;
        VecList vl;
        for (int i = 0; i < KSFD_MAXV; ++i)
            vl.v[i] = V + (long long)(b + std::min(i, m - 1)) * n;
        const double *inv = last ? aux : nullptr;
        switch (m) {
        case 1: TRY(orth_scale_launch<1>(c, vl, hdev + b, inv, w, st)); break;
        case 2: TRY(orth_scale_launch<2>(c, vl, hdev + b, inv, w, st)); break;
        case 3: TRY(orth_scale_launch<3>(c, vl, hdev + b, inv, w, st)); break;
        case 4: TRY(orth_scale_launch<4>(c, vl, hdev + b, inv, w, st)); break;
        case 5: TRY(orth_scale_launch<5>(c, vl, hdev + b, inv, w, st)); break;
        case 6: TRY(orth_scale_launch<6>(c, vl, hdev + b, inv, w, st)); break;
        case 7: TRY(orth_scale_launch<7>(c, vl, hdev + b, inv, w, st)); break;
        default: TRY(orth_scale_launch<8>(c, vl, hdev + b, inv, w, st)); break;
        }
    }
    return 0;
}

static int gmres_sync_impl(ksfd_ctx *c, const double *rhs, double rhs_sign, double *x,
                           const ksfd_ksp_opts &o, ksfd_ksp_result *res, cudaStream_t st)
{
    const long long n = nlocal(c);
    if (o.restart > KSFD_MAX_RESTART)
        return fail("ksfd_gmres: restart " + std::to_string(o.restart) + " exceeds the supported maximum of " +
                    std::to_string(KSFD_MAX_RESTART) + " (the device-side Hessenberg lives in shared memory)");
    const int m = std::max(1, o.restart > 0 ? o.restart : 30);
    const int max_it = o.max_it > 0 ? o.max_it : 10000;
    const bool pre = o.precond != 0;
    if (c->krylov_cap < m + 1) {
        cudaFree(c->krylov);
        c->krylov = nullptr;
        CK(cudaMalloc(&c->krylov, sizeof(double) * n * (m + 1)));
        c->krylov_cap = m + 1;
    }
    TRY(ensure_work(c, 0));
    double *V = c->krylov;
    double *tmp = c->work[0];
    double *hdev = c->dscal + SC_H;   // Hessenberg column j: hdev[0..j+1]
    std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), gg(m + 1), y(m);
    CK(cudaMemsetAsync(x, 0, sizeof(double) * n, st));
    int its = 0, reason = 0;
    double rnorm0 = 0.0, rnorm = 0.0, tol = 0.0;
    bool first = true;
    while (true) {
        const double *r;
        double sign;
        if (first) {
            r = rhs;
            sign = rhs_sign;
        } else {
            // r = sign*rhs - A x   (true residual at restart)
            TRY(jvp_impl(c, x, tmp, false, st));
            CoefList cl{};
            VecList vl{};
            cl.c[0] = rhs_sign;
            vl.v[0] = rhs;
            k_maxpy_host<1><<<KSFD_RED_BLOCKS, 256, 0, st>>>(n, cl, vl, -1.0, tmp);
            CKL();
            r = tmp;
            sign = 1.0;
        }
        TRY(norm2_dev(c, r, SC_NORM, st));
        TRY(fetch(c, SC_NORM, 1, st));
        const double beta = c->hscal[SC_NORM];
        if (first) {
            rnorm0 = beta;
            tol = std::max(o.rtol * beta, o.atol);
            first = false;
        }
        rnorm = beta;
        if (!(beta == beta)) { reason = -9; break; }          // NaN
        if (beta <= tol || beta == 0.0) { reason = beta == 0.0 ? 3 : 2; break; }
        if (its >= max_it) { reason = -3; break; }
        k_scale_by_inv<<<KSFD_RED_BLOCKS, 256, 0, st>>>(n, r, c->dscal + SC_NORM, sign, V);
        CKL();
        gg.assign(m + 1, 0.0);
        gg[0] = beta;
        int j = 0;
        bool done = false;
        // Single-pass classical Gram-Schmidt loses orthogonality like
        // eps*kappa^2, kappa ~ (residual reduction inside the cycle), which
        // makes the recurrence residual unreliable beyond ~1e-6.  Instead of
        // paying for a second orthogonalisation pass every iteration, a cycle
        // is closed (x updated, TRUE residual recomputed) once it has reduced
        // the residual by 1e-5; the next cycle starts from exact data.
        const double cycle_tol = o.reorth ? tol : std::max(tol, 1e-5 * beta);
        for (; j < m; ++j) {
            double *w = V + (long long)(j + 1) * n;
            TRY(jvp_impl(c, V + (long long)j * n, w, pre, st));
            double *Hc = &H[(size_t)j * (m + 1)];
            bool explicit_norm = o.reorth != 0;
            if (!o.reorth) {
                // fused single-pass CGS: 4 launches, norm from <w,w> - |h|^2
                TRY(cgs_fused_step(c, j, V, w, st));
                TRY(fetch(c, SC_H, 128, st));      // column + aux in one copy
                for (int i = 0; i <= j + 1; ++i) Hc[i] = c->hscal[SC_H + i];
                if (c->hscal[SC_AUX + 1] != 0.0) {
                    // heavy cancellation: w was orthogonalised but left unscaled;
                    // take its norm explicitly
                    TRY(norm2_dev(c, w, SC_H + j + 1, st));
                    TRY(fetch(c, SC_H + j + 1, 1, st));
                    Hc[j + 1] = c->hscal[SC_H + j + 1];
                    explicit_norm = true;
                }
            } else {
                std::vector<const double *> vp(j + 1);
                for (int i = 0; i <= j; ++i) vp[i] = V + (long long)i * n;
                TRY(mdot_impl(c, j + 1, vp.data(), w, hdev, st));
                TRY(orth_impl(c, j + 1, V, hdev, w, hdev + j + 1, st));
                // second classical Gram-Schmidt pass (CGS2)
                TRY(mdot_impl(c, j + 1, vp.data(), w, c->dscal + SC_H2, st));
                TRY(orth_impl(c, j + 1, V, c->dscal + SC_H2, w, hdev + j + 1, st));
                CK(cudaMemcpyAsync(c->hscal + SC_H2, c->dscal + SC_H2,
                                   sizeof(double) * (j + 1), cudaMemcpyDeviceToHost, st));
                TRY(fetch(c, SC_H, j + 2, st));
                for (int i = 0; i <= j + 1; ++i) Hc[i] = c->hscal[SC_H + i];
                for (int i = 0; i <= j; ++i) Hc[i] += c->hscal[SC_H2 + i];
            }
            const double hnext = Hc[j + 1];
            if (explicit_norm && hnext > 0.0) {
                k_scale_by_inv<<<KSFD_RED_BLOCKS, 256, 0, st>>>(n, w, hdev + j + 1, 1.0, w);
                CKL();
            }
            for (int i = 0; i < j; ++i) {
                const double t = cs[i] * Hc[i] + sn[i] * Hc[i + 1];
                Hc[i + 1] = -sn[i] * Hc[i] + cs[i] * Hc[i + 1];
                Hc[i] = t;
            }
            const double den = std::hypot(Hc[j], Hc[j + 1]);
            cs[j] = den > 0 ? Hc[j] / den : 1.0;
            sn[j] = den > 0 ? Hc[j + 1] / den : 0.0;
            Hc[j] = den;
            Hc[j + 1] = 0.0;
            gg[j + 1] = -sn[j] * gg[j];
            gg[j] = cs[j] * gg[j];
            rnorm = std::fabs(gg[j + 1]);
            ++its;
            if (!(rnorm == rnorm)) { reason = -9; done = true; ++j; break; }
            if (rnorm <= tol && cycle_tol <= tol) { reason = 2; done = true; ++j; break; }
            if (o.dtol > 0 && rnorm > o.dtol * rnorm0) { reason = -4; done = true; ++j; break; }
            if (its >= max_it) { reason = -3; done = true; ++j; break; }
            if (hnext == 0.0) { reason = 2; done = true; ++j; break; }
            if (rnorm <= cycle_tol) { ++j; break; }       // close the cycle, re-verify
        }
        // back substitution on the j x j triangle
        const int k = j;
        for (int i = k - 1; i >= 0; --i) {
            double s = gg[i];
            for (int q = i + 1; q < k; ++q) s -= H[(size_t)q * (m + 1) + i] * y[q];
            y[i] = s / H[(size_t)i * (m + 1) + i];
        }
        if (reason != -9 && k > 0) {
            std::vector<const double *> vp(k);
            for (int i = 0; i < k; ++i) vp[i] = V + (long long)i * n;
            if (pre) {
                // x += M^{-1} (V y)
                TRY(ensure_work(c, 1));
                TRY(maxpy_impl(c, k, y.data(), vp.data(), 0.0, tmp, st));
                TRY(pc_apply_impl(c, tmp, c->work[1], st));
                const double one = 1.0;
                const double *wp = c->work[1];
                TRY(maxpy_impl(c, 1, &one, &wp, 1.0, x, st));
            } else {
                TRY(maxpy_impl(c, k, y.data(), vp.data(), 1.0, x, st));
            }
        }
        if (done) break;
    }
    if (res) {
        res->its = its;
        res->reason = reason;
        res->rnorm0 = rnorm0;
        res->rnorm = rnorm;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Pipelined GMRES(m) (default): same algorithm as gmres_sync_impl (single-pass
// classical Gram-Schmidt with the norm from <w,w> - |h|^2, cycles closed after
// a 1e-5 reduction and restarted from the TRUE residual), but the Hessenberg /
// Givens bookkeeping, the convergence tests and the back substitution run on
// the device.  The host never synchronises the stream inside a solve: it polls
// a pinned status word and launches up to `gm_runahead` Arnoldi steps ahead;
// kernels launched past the end of a cycle return at once (skip flags).
// With several ranks the run-ahead is 0, so that every rank issues the same
// sequence of NCCL calls (decisions are taken from all-reduced, identical data).
// ---------------------------------------------------------------------------
static int gm_alloc(ksfd_ctx *c)
{
    if (c->gm) return 0;
    CK(cudaMalloc(&c->gm, sizeof(double) * GM_DOUBLES));
    CK(cudaMalloc(&c->gmi, sizeof(int) * GMI_INTS));
    CK(cudaMalloc(&c->gm_done, sizeof(unsigned)));
    CK(cudaMemset(c->gm_done, 0, sizeof(unsigned)));
    CK(cudaHostAlloc(&c->gm_status, sizeof(GmStatus), cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&c->gm_status_dev, c->gm_status, 0));
    memset(c->gm_status, 0, sizeof(GmStatus));
    return 0;
}

// spin until pred() holds; fails if the stream dies or drains without it
template <class Pred>
static int gm_wait(cudaStream_t st, Pred pred, const char *what, const ksfd_ctx *c = nullptr)
{
    time_t t0 = 0;
    bool warned = false;
    for (unsigned long long spins = 0;; ++spins) {
        if (pred()) return 0;
        if ((spins & 0x3ff) == 0x3ff) {
            if (c && c->p2p_err && *c->p2p_err)
                return fail("peer-to-peer exchange timed out (a neighbouring rank is gone)");
            if (!t0) t0 = time(nullptr);
            if (!warned && time(nullptr) - t0 > 20) {
                warned = true;
                const GmStatus *hs = c ? static_cast<const GmStatus *>(c->gm_status) : nullptr;
                fprintf(stderr, "ksfd_b200[rank %d]: GMRES pipeline waiting >20 s for %s "
                        "(seq %d iters %d cycle_done %d final %d its %d)\n",
                        c ? c->rank : -1, what, hs ? hs->seq : -1, hs ? hs->iters_done : -1,
                        hs ? hs->cycle_done : -1, hs ? hs->final_ : -1, hs ? hs->its_total : -1);
            }
            cudaError_t e = cudaStreamQuery(st);
            if (e == cudaSuccess) {
                if (pred()) return 0;
                return fail(std::string("pipelined GMRES stalled waiting for ") + what);
            }
            if (e != cudaErrorNotReady)
                return fail(std::string("pipelined GMRES: ") + cudaGetErrorString(e));
        }
    }
}

template <int NV>
static int gm_mdot_launch(ksfd_ctx *c, const VecList &vl, const double *w, const GmFin *fin,
                          cudaStream_t st)
{
    ProfScope prof(c, 2, st);
    if (fin) {
        KSFD_KLAUNCH((k_gm_mdot<NV, true>), KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st, nlocal(c), vl, w,
                     c->gmi, c->partial, *fin);
    } else {
        GmFin none{};
        KSFD_KLAUNCH((k_gm_mdot<NV, false>), KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st, nlocal(c), vl, w,
                     c->gmi, c->partial, none);
    }
    CKL();
    return 0;
}
template <int NV>
static int gm_orth_launch(ksfd_ctx *c, const VecList &vl, int off, int do_scale, double *w,
                          cudaStream_t st)
{
    // the launch that finishes the new basis vector also pushes its boundary planes to
    // the neighbours (the vector is the operand of the next J.v)
    ProfScope prof(c, 3, st);
    HaloPush hp{};
    if (do_scale) {
        hp = make_push(c, 1);
        if (hp.up_lo0) c->pushed_vec = w;
    }
    if (hp.up_lo0)
        KSFD_KLAUNCH((k_gm_orth_scale<NV, true>), KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st, nlocal(c), vl,
                     off, do_scale, c->gm, c->gmi, w, hp);
    else
        KSFD_KLAUNCH((k_gm_orth_scale<NV, false>), KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st, nlocal(c), vl,
                     off, do_scale, c->gm, c->gmi, w, hp);
    CKL();
    return 0;
}
#define GM_SWITCH(m_, CALL)                      \
    switch (m_) {                                \
    case 1: TRY(CALL(1)); break;                 \
    case 2: TRY(CALL(2)); break;                 \
    case 3: TRY(CALL(3)); break;                 \
    case 4: TRY(CALL(4)); break;                 \
    case 5: TRY(CALL(5)); break;                 \
    case 6: TRY(CALL(6)); break;                 \
    case 7: TRY(CALL(7)); break;                 \
    default: TRY(CALL(8)); break;                \
    }

// one Arnoldi step j of the pipeline (all launches, no host wait)
static int gm_step(ksfd_ctx *c, int j, double *V, int pcm, const GmOpts &go, cudaStream_t st)
{
    const long long n = nlocal(c);
    double *w = V + (long long)(j + 1) * n;
    GmStatus *hsd = static_cast<GmStatus *>(c->gm_status_dev);
    if (pcm == 2) {
        // w = A (A0^-1 v_j): spectral preconditioner, then the plain stencil pass
        TRY(fftpc_apply(c, V + (long long)j * n, c->work[11], c->work[10], st,
                        c->gmi + GMI_CYCLE_DONE));
        TRY(jvp_impl(c, c->work[10], w, false, st, c->gmi + GMI_CYCLE_DONE));
    } else {
        TRY(jvp_impl(c, V + (long long)j * n, w, pcm == 1, st, c->gmi + GMI_CYCLE_DONE));
    }
    const int nv = j + 2;                     // V_0..V_j and w itself
    for (int b = 0; b < nv; b += KSFD_MAXV) {
        const int m = std::min(KSFD_MAXV, nv - b);
        const bool last = b + m >= nv;
        VecList vl;
        for (int i = 0; i < KSFD_MAXV; ++i) {
            const int idx = b + std::min(i, m - 1);
            vl.v[i] = (idx == j + 1) ? w : V + (long long)idx * n;
        }
        const bool fused = c->nranks == 1 || c->p2p_on;     // rank sum inside the kernel
        GmFin fin{m, b, j, KSFD_RED_BLOCKS, c->partial, c->gm, c->gmi, hsd, go, p2p_red(c),
                  c->gm_done};
        const GmFin *tail = (last && fused) ? &fin : nullptr;   // bookkeeping in the last block
#define GM_CALL_MDOT(N) gm_mdot_launch<N>(c, vl, w, tail, st)
        GM_SWITCH(m, GM_CALL_MDOT)
#undef GM_CALL_MDOT
        if (!tail) {
            k_gm_reduce<<<m, 128, 0, st>>>(m, b, KSFD_RED_BLOCKS, c->partial, c->gmi, c->gm);
            CKL();
        }
    }
    if (c->nranks > 1 && !c->p2p_on) {
        P2PRed none{};
        none.nranks = 1;
        TRY(allreduce_dev(c, c->gm + GM_HCOL, nv, ncclSum_, st));
        GmFin fin{0, 0, j, 0, c->partial, c->gm, c->gmi, hsd, go, none, c->gm_done};
        k_gm_finalize<<<1, 256, 0, st>>>(fin);
        CKL();
    }
    const int no = j + 1;
    for (int b = 0; b < no; b += KSFD_MAXV) {
        const int m = std::min(KSFD_MAXV, no - b);
        const bool last = b + m >= no;
        VecList vl;
        for (int i = 0; i < KSFD_MAXV; ++i)
            vl.v[i] = V + (long long)(b + std::min(i, m - 1)) * n;
#define GM_CALL_ORTH(N) gm_orth_launch<N>(c, vl, b, last ? 1 : 0, w, st)
        GM_SWITCH(m, GM_CALL_ORTH)
#undef GM_CALL_ORTH
    }
    return 0;
}

// cycle0 = 1: continue a solve the Richardson sweeps started (sweep_solve_impl): x holds their
// iterate, the device state their tolerance, ||b|| and iteration count; the first cycle
// starts from the true residual of that x
static int gmres_pipe_impl(ksfd_ctx *c, const double *rhs, double rhs_sign, double *x,
                           const ksfd_ksp_opts &o, ksfd_ksp_result *res, cudaStream_t st,
                           int cycle0 = 0)
{
    const long long n = nlocal(c);
    if (o.restart > KSFD_MAX_RESTART)
        return fail("ksfd_gmres: restart " + std::to_string(o.restart) + " exceeds the supported maximum of " +
                    std::to_string(KSFD_MAX_RESTART) + " (the device-side Hessenberg lives in shared memory)");
    const int m = std::max(1, o.restart > 0 ? o.restart : 30);
    // preconditioner: 0 none, 1 point-block Jacobi (fused into the stencil
    // kernel), 2 spectral (fftpc.cuh), 3 automatic: block Jacobi until a solve
    // needs >= 16 steps, then spectral; back after three spectral solves of <= 2
    // steps.  Spectral needs one rank and cuFFT, else block Jacobi.
    int pcm = o.precond;
    if (pcm == 3) pcm = c->pc_auto_fft ? 2 : 1;
    if (pcm == 2 && !fftpc_prepare(c)) pcm = 1;
    const bool pre = pcm == 1;
    if (c->krylov_cap < m + 1) {
        cudaFree(c->krylov);
        c->krylov = nullptr;
        CK(cudaMalloc(&c->krylov, sizeof(double) * n * (m + 1)));
        c->krylov_cap = m + 1;
    }
    TRY(ensure_work(c, 0));
    TRY(gm_alloc(c));
    if (pcm == 2) {
        TRY(ensure_work(c, 10));
        TRY(ensure_work(c, 11));
        if (!c->fft_means_valid) {
            TRY(fftpc_setup(c, st));
            c->fft_means_valid = true;
        }
    }
    double *V = c->krylov;
    double *tmp = c->work[0];
    GmStatus *hs = static_cast<GmStatus *>(c->gm_status);
    GmStatus *hsd = static_cast<GmStatus *>(c->gm_status_dev);
    GmOpts go{o.rtol, o.atol, o.dtol, o.max_it > 0 ? o.max_it : 10000, m, 0, c->gm_cycle_factor};
    const int R = c->gm_runahead;
    // with the peer-to-peer exchanges (device-side exchange counters) launch
    // decisions need not be identical on all ranks
    // (the slab-distributed spectral preconditioner makes NCCL calls in every step:
    // identical launch sequences on all ranks are required, as in the NCCL fallback)
    const bool free_running = c->nranks == 1 || (c->p2p_on && !(pcm == 2 && c->nranks > 1));
    InvD id;
    for (int l = 0; l < KSFD_MAX_LIGANDS; ++l) id.v[l] = c->invd[l];

    c->pushed_vec = nullptr;
    c->gm_no_push = pcm == 2;       // the spectral preconditioner sits between V_j and the stencil pass
    // the previous solve on this context has been waited for (see the end of
    // this function), so the status block is ours
    hs->seq = 0;
    hs->iters_done = 0;
    hs->cycle_done = hs->final_ = hs->reason = hs->its_total = 0;
    hs->k_cols = 0;
    // (continuing: only the skip flags GMI_CYCLE_DONE, GMI_FINAL, GMI_NOUPD are cleared)
    CK(cudaMemsetAsync(c->gmi, 0, sizeof(int) * (cycle0 ? 3 : GMI_INTS), st));
    // x = 0 is not stored: the first cycle's update WRITES x (x_zero below)
    for (int cycle = cycle0;; ++cycle) {
        const double *r = rhs;
        double sign = rhs_sign;
        // one rank: the block that finishes the <r,r> reduction last also starts
        // the cycle (no separate one-block launch)
        const bool fuse_begin = c->nranks == 1 || c->p2p_on;
        GmBegin gb{KSFD_RED_BLOCKS, c->partial, c->gm, c->gmi, hsd, cycle, go, p2p_red(c),
                   c->gm_done};
        {
        ProfScope prof_begin(c, 5, st);
        if (cycle == 0) {
            if (fuse_begin) {
                KSFD_KLAUNCH(k_gm_norm_begin, KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st, n, rhs,
                             c->partial, gb);
            } else {
                VecList vl;
                for (int i = 0; i < KSFD_MAXV; ++i) vl.v[i] = rhs;
                k_mdot<1><<<KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st>>>(n, vl, rhs, c->partial);
            }
            CKL();
        } else {
            // r = sign*rhs - A x   (true residual at restart)
            TRY(jvp_impl(c, x, tmp, false, st, c->gmi + GMI_FINAL));
            if (fuse_begin)
                KSFD_KLAUNCH((k_gm_true_residual<true>), KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st, n,
                             rhs, rhs_sign, c->gmi, tmp, c->partial, gb);
            else
                KSFD_KLAUNCH((k_gm_true_residual<false>), KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st, n,
                             rhs, rhs_sign, c->gmi, tmp, c->partial, gb);
            CKL();
            r = tmp;
            sign = 1.0;
        }
        if (fuse_begin) {
            // done above
        } else if (c->p2p_on) {
            k_gm_cycle_begin<<<1, 256, 0, st>>>(gb);
            CKL();
        } else {
            P2PRed none{};
            none.nranks = 1;
            k_reduce_partials<<<1, 128, 0, st>>>(1, KSFD_RED_BLOCKS, c->partial,
                                                 c->dscal + SC_NORM, 0);
            CKL();
            TRY(allreduce_dev(c, c->dscal + SC_NORM, 1, ncclSum_, st));
            GmBegin g1{1, c->dscal + SC_NORM, c->gm, c->gmi, hsd, cycle, go, none, c->gm_done};
            k_gm_cycle_begin<<<1, 32, 0, st>>>(g1);
            CKL();
        }
        }
        {
            ProfScope prof(c, 4, st);
            const HaloPush hp = make_push(c, 1);
            if (hp.up_lo0) c->pushed_vec = V;
            if (hp.up_lo0)
                KSFD_KLAUNCH(k_gm_first_vector<true>, KSFD_RED_BLOCKS, 256, 0, st, n, r, c->gm, c->gmi,
                             sign, V, hp);
            else
                KSFD_KLAUNCH(k_gm_first_vector<false>, KSFD_RED_BLOCKS, 256, 0, st, n, r, c->gm, c->gmi,
                             sign, V, hp);
            CKL();
        }
        const int seq = 2 * cycle + 1;
        if (free_running) {
            for (int j = 0; j < m; ++j) {
                // launch step j once step j-R-1 is known not to have closed the
                // cycle; the first R+1 steps go out before the cycle has even
                // begun on the device (they skip if it ends at once)
                // cycles of consecutive solves have very similar lengths: launch
                // ahead only while the cycle is expected to go on, so that few
                // launched-ahead steps are wasted (each costs ~5 skipped kernels)
                const int pred = c->gm_pred[cycle > 0];
                const int Rj = (pred == 0 || j + 1 < pred) ? R : 0;
                TRY(gm_wait(st, [&] {
                    if (hs->seq < seq) return j <= Rj;
                    return hs->cycle_done != 0 || hs->iters_done >= j - Rj;
                }, "an Arnoldi step", c));
                if (hs->seq >= seq && hs->cycle_done) break;
                TRY(gm_step(c, j, V, pcm, go, st));
            }
        } else {
            // several ranks over NCCL: every launch decision must be taken from the
            // same data on all ranks, or their NCCL call sequences diverge.  Steps go
            // out in chunks of `chunk`; the decision to launch the next chunk is
            // taken only when the previous one has completely finished (its
            // status is final and identical everywhere); steps of a chunk past
            // the end of the cycle skip their kernels but still make their
            // (matching) NCCL calls.
            const int chunk = std::max(1, c->gm_runahead);
            for (int j = 0; j < m; j += chunk) {
                TRY(gm_wait(st, [&] {
                    return hs->seq >= seq && (hs->cycle_done != 0 || hs->iters_done >= j);
                }, "an Arnoldi chunk", c));
                if (hs->cycle_done) break;
                for (int jj = j; jj < std::min(j + chunk, m); ++jj)
                    TRY(gm_step(c, jj, V, pcm, go, st));
            }
        }
        TRY(gm_wait(st, [&] { return hs->seq >= seq && hs->cycle_done; }, "the end of a cycle", c));
        if (getenv("KSFD_DEBUG_GMRES"))
            fprintf(stderr, "  cycle %d: seq %d iters %d cycle_done %d final %d reason %d its %d "
                    "rnorm %.3e\n", cycle, hs->seq, hs->k_cols, hs->cycle_done, hs->final_,
                    hs->reason, hs->its_total, hs->rnorm);
        c->gm_pred[cycle > 0] = hs->k_cols;
        if (hs->reason != -9 && hs->k_cols > 0) {
            const unsigned ub = std::min(nblk(c->g.npts, 256), 148u * 8u);
#define KSFD_UPD(D, PCF, VV, YY, KK, XZ, XX)                                                 \
    KSFD_KLAUNCH((k_gm_update_x<D>), ub, 256, 0, st, c->g, c->Pjac, coef_ref(c), id, c->pc, PCF, n, VV, \
                 YY, KK, c->gmi + GMI_NOUPD, XZ, XX)
#define KSFD_UPD_DOF(PCF, VV, YY, KK, XZ, XX)                                                \
    switch (c->dof) {                                                                        \
    case 2: KSFD_UPD(2, PCF, VV, YY, KK, XZ, XX); break;                                     \
    case 3: KSFD_UPD(3, PCF, VV, YY, KK, XZ, XX); break;                                     \
    case 4: KSFD_UPD(4, PCF, VV, YY, KK, XZ, XX); break;                                     \
    case 5: KSFD_UPD(5, PCF, VV, YY, KK, XZ, XX); break;                                     \
    default: KSFD_UPD(0, PCF, VV, YY, KK, XZ, XX); break;                                    \
    }
            const int xz = cycle == 0 ? 1 : 0;
            if (pcm == 2) {
                // x += A0^-1 (V y)
                double *z = c->work[10], *s2 = c->work[11];
                KSFD_UPD_DOF(0, V, c->gm + GM_Y, c->gmi + GMI_K, 1, z)
                CKL();
                TRY(fftpc_apply(c, z, z, s2, st, c->gmi + GMI_NOUPD));
                KSFD_UPD_DOF(0, s2, nullptr, nullptr, xz, x)
            } else {
                KSFD_UPD_DOF(pre ? 1 : 0, V, c->gm + GM_Y, c->gmi + GMI_K, xz, x)
            }
#undef KSFD_UPD_DOF
#undef KSFD_UPD
            CKL();
        } else if (cycle == 0) {
            CK(cudaMemsetAsync(x, 0, sizeof(double) * n, st));     // nothing to add: x = 0
        }
        if (hs->final_) break;
    }
    if (getenv("KSFD_DEBUG_GMRES"))
        fprintf(stderr, "gmres: its %d reason %d rnorm0 %.3e rnorm %.3e\n", hs->its_total,
                hs->reason, hs->rnorm0, hs->rnorm);
    if (o.precond == 3 && hs->reason > 0) {
        if (pcm == 1 && hs->its_total >= 16 && fftpc_available(c)) {
            c->pc_auto_fft = true;
            c->pc_auto_small = 0;
        } else if (pcm == 2) {
            c->pc_auto_small = hs->its_total <= 2 ? c->pc_auto_small + 1 : 0;
            if (c->pc_auto_small >= 3) c->pc_auto_fft = false;
        }
    }
    if (res) {
        res->its = hs->its_total;
        res->reason = hs->reason;
        res->rnorm0 = hs->rnorm0;
        res->rnorm = hs->rnorm;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Preconditioned Richardson sweeps (sweep_op.cuh): one fused stencil pass per iteration,
// convergence decided on the device, the host launches ahead and polls the pinned
// status block exactly as for the pipelined GMRES.  Needs the marching kernels, the fused
// block-Jacobi preconditioner and (several ranks) the peer-memory exchange.
// ---------------------------------------------------------------------------
#define KSFD_SWEEP_CTAS 262144         // CTAs of a sweep grid the partial-sum buffer (3 x 2 MB) has room for
static bool sweep_eligible(const ksfd_ctx *c, const ksfd_ksp_opts &o)
{
    if (o.ksp_type == 0 || o.reorth || !c->gm_pipeline) return false;
    if (!use_march(c) || o.precond == 0 || o.precond == 2) return false;
    if (o.precond == 3 && c->pc_auto_fft) return false;    // large steps: spectral + GMRES
    if (c->nranks > 1 && !c->p2p_on) return false;
    return true;
}

// State of one solve by sweeps between its two halves: sweep_begin launches the predicted
// sweeps and returns; the caller may enqueue the work that FOLLOWS the solve (the next
// stage's combination and residual) before sweep_end waits for the decision, so the device
// does not idle during the host round trip at the end of a solve.  `clean` tells the caller
// whether that work saw the final solution (the prediction held) or must be redone.
struct SweepRun {
    const double *rhs;
    double rhs_sign;
    double *x;
    ksfd_ksp_opts o;
    cudaStream_t st;
    double *buf[2];
    GmStatus *hs, *hsd;
    GmOpts go;
    HaloPush hp;
    HostVec ph, ch;
    bool pure, defer, fuse_push;
    // several ranks: work enqueued behind the predicted sweeps has made exchanges on the halo
    // slot of the sweeps (the next right-hand side was pushed): a sweep launched after it
    // must push its input again
    bool dirty;
    int slot, pred, lead, launched, first;
    const int *skip;
};

static int sweep_launch(ksfd_ctx *c, SweepRun &r, int it)
{
    cudaStream_t st = r.st;
    const double *rin = it == 0 ? r.rhs : r.buf[(it - 1) & 1];
    if (it == 0 && r.fuse_push && c->pushed_vec == r.rhs)
        c->pushed_vec = nullptr;            // pushed by the residual kernel that produced it
    else if (it == 0 || !r.fuse_push)
        TRY(exchange(c, rin, c->dof, 1, st, r.skip, r.defer));
    const HostVec rh = make_hvec(c, rin, c->dof, 1);
    // sweeps the solve is known to need are not tested (no reduction, no rank sum)
    // (the last sweep that goes out without waiting is always tested: the host waits for it)
    const int test = (r.pred == 0 || it >= r.pred - r.lead || it >= r.first - 1) ? 1 : 0;
    SweepFin fin{c->sw_partial, KSFD_SWEEP_CTAS, it, test, 0, c->gm, c->gmi, r.hsd, r.go,
                 r.pure ? 1e300 : c->sw_slow, p2p_red(c), c->gm_done};
    SweepHost a{r.x, r.buf[it & 1], it == 0 ? r.rhs_sign : 1.0, it == 0 ? 1 : 0, KSFD_SWEEP_CTAS,
                &fin, r.fuse_push ? &r.hp : nullptr};
    ProfScope prof(c, 6, st);
    return c->dim == 2 ? ksfd_march_sweep_d2(c, r.ch, rh, r.ph, a, r.skip, st)
                       : ksfd_march_sweep_d3(c, r.ch, rh, r.ph, a, r.skip, st);
}

static int sweep_begin(ksfd_ctx *c, SweepRun &r, const double *rhs, double rhs_sign, double *x,
                       const ksfd_ksp_opts &o, cudaStream_t st)
{
    const long long n = nlocal(c);
    const int m = std::max(1, o.restart > 0 ? o.restart : 30);
    if (c->krylov_cap < m + 1) {            // sized as the GMRES fallback wants it
        cudaFree(c->krylov);
        c->krylov = nullptr;
        CK(cudaMalloc(&c->krylov, sizeof(double) * n * (m + 1)));
        c->krylov_cap = m + 1;
    }
    TRY(gm_alloc(c));
    if (!c->sw_partial) CK(cudaMalloc(&c->sw_partial, sizeof(double) * 3 * KSFD_SWEEP_CTAS));
    r.rhs = rhs;
    r.rhs_sign = rhs_sign;
    r.x = x;
    r.o = o;
    r.st = st;
    r.buf[0] = c->krylov;
    r.buf[1] = c->krylov + n;
    r.hs = static_cast<GmStatus *>(c->gm_status);
    r.hsd = static_cast<GmStatus *>(c->gm_status_dev);
    r.pure = o.ksp_type == 1;               // no fallback: stop on max_it / dtol only
    r.go = GmOpts{o.rtol, o.atol, o.dtol, o.max_it > 0 ? o.max_it : 10000, m, 0, 0.0};
    GmStatus *hs = r.hs;
    hs->seq = 0;
    hs->iters_done = 0;
    hs->cycle_done = hs->final_ = hs->reason = hs->its_total = 0;
    hs->k_cols = 0;
    CK(cudaMemsetAsync(c->gmi, 0, sizeof(int) * GMI_INTS, st));
    r.skip = c->gmi + GMI_CYCLE_DONE;
    r.defer = c->p2p_on && c->nranks > 1 && tma_consumer(c);
    r.ph = make_hvec(c, c->pc, 1, 2);
    r.ch = coef_hvec(c);
    // several ranks: every sweep also pushes the boundary planes of its output into the
    // neighbours' ghost buffers (no exchange kernel between sweeps); only a right-hand
    // side nobody pushed goes out through the push kernel
    r.hp = make_push(c, 1);
    r.fuse_push = r.defer && r.hp.up_lo0 != nullptr && ksfd_use_tma(c) && c->sw_fuse_push;
    r.slot = c->sw_slot & 3;
    r.pred = c->sw_hist[r.slot];
    // one more tested sweep while the count of this slot is not settled, and every 8th solve
    const bool probe = c->sw_stable[r.slot] < 2 || c->sw_stable[r.slot] % 8 == 7;
    r.lead = c->sw_test_lead + (probe ? 1 : 0);
    // Solves of the same slot (ROSW stage) of consecutive steps take the same number of
    // sweeps.  The predicted number goes out without waiting and only its last sweep (while
    // the count is not settled, and every 8th solve: its last two) is tested; beyond it the
    // host launches one tested sweep at a time, each once the previous one is known not to
    // have ended the solve (a host round trip, ~10 us: the price of a prediction that was too
    // short).  Without a prediction every sweep is tested and the host stays R sweeps ahead.
    r.launched = 0;
    r.dirty = false;
    // (test knob sweep_underpredict: launch fewer than predicted, to exercise the late path)
    r.first = std::max(0, std::min(r.pred, r.go.max_it) - (r.pred > 1 ? c->sw_underpredict : 0));
    for (; r.launched < r.first; ++r.launched) TRY(sweep_launch(c, r, r.launched));
    return 0;
}

// the prediction of this solve is settled: work enqueued behind the predicted sweeps will
// most probably see the converged solution
static bool sweep_settled(const ksfd_ctx *c, const SweepRun &r)
{
    return r.pred > 0 && r.launched >= r.pred - c->sw_underpredict && r.launched > 0 &&
           c->sw_stable[r.slot] >= 2;
}

static int sweep_end(ksfd_ctx *c, SweepRun &r, ksfd_ksp_result *res, bool *clean)
{
    cudaStream_t st = r.st;
    GmStatus *hs = r.hs;
    const int R = c->gm_runahead;
    const int launched0 = r.launched;
    for (;;) {
        const int lag = r.pred > 0 ? 0 : R;
        TRY(gm_wait(st, [&] { return hs->cycle_done != 0 || hs->iters_done >= r.launched - lag; },
                    "a Richardson sweep", c));
        // The sweep that ends a solve writes final_, a system fence, cycle_done and then
        // iters_done (sweep_finalize); the last two are not ordered against each other, so
        // iters_done may be seen first.  final_ is older than the fence: whoever sees that
        // iters_done also sees it, and must not launch one more sweep.  (On several ranks an
        // extra sweep on one rank alone would also make that rank redo the speculated work and
        // its halo pushes: exchange numbers would run apart.)  The fence keeps the loads below
        // behind the ones of the wait on hosts that reorder loads.
        std::atomic_thread_fence(std::memory_order_acquire);
        if (hs->cycle_done || hs->final_ || r.launched >= r.go.max_it) break;
        if (r.dirty && r.launched > 0) {
            TRY(exchange(c, r.buf[(r.launched - 1) & 1], c->dof, 1, st, r.skip, r.defer));
            r.dirty = false;
        }
        TRY(sweep_launch(c, r, r.launched));
        ++r.launched;
    }
    // iters_done is the LAST word a solve writes into the status block: the block is reset for
    // the next solve (sweep_begin) only once it has landed, or the late store would be taken
    // for a finished sweep of that solve (sweeps launched beyond the end skip and write nothing)
    TRY(gm_wait(st, [&] { return hs->cycle_done != 0 && hs->iters_done == hs->its_total; },
                "the end of the sweeps", c));
    if (getenv("KSFD_DEBUG_GMRES"))
        fprintf(stderr, "sweeps: its %d reason %d rnorm0 %.3e rnorm %.3e\n", hs->its_total,
                hs->reason, hs->rnorm0, hs->rnorm);
    if (clean) *clean = r.launched == launched0 && hs->reason > 0 && hs->reason != KSFD_SWEEP_FALLBACK;
    if (hs->reason == KSFD_SWEEP_FALLBACK) {
        // contraction too slow for a stationary iteration: GMRES takes over, from the
        // iterate reached so far unless it is worse than x = 0
        c->sw_backoff = 8;
        for (int i = 0; i < 4; ++i) c->sw_hist[i] = c->sw_stable[i] = 0;
        const bool keep = hs->rnorm < hs->rnorm0;
        return gmres_pipe_impl(c, r.rhs, r.rhs_sign, r.x, r.o, res, st, keep ? 1 : 0);
    }
    c->sw_stable[r.slot] = (r.pred > 0 && hs->its_total == r.pred) ? c->sw_stable[r.slot] + 1 : 0;
    c->sw_hist[r.slot] = hs->its_total;
    if (res) {
        res->its = hs->its_total;
        res->reason = hs->reason;
        res->rnorm0 = hs->rnorm0;
        res->rnorm = hs->rnorm;
    }
    return 0;
}

static int sweep_solve_impl(ksfd_ctx *c, const double *rhs, double rhs_sign, double *x,
                            const ksfd_ksp_opts &o, ksfd_ksp_result *res, cudaStream_t st)
{
    SweepRun r;
    TRY(sweep_begin(c, r, rhs, rhs_sign, x, o, st));
    return sweep_end(c, r, res, nullptr);
}

// one sweep as a stand-alone operation (tests, kernel timing): r_out = r_in - A M^-1 r_in,
// x = (first ? 0 : x) + M^-1 r_in, norms[0] = ||r_in||, norms[1] = ||r_out|| (all ranks)
extern "C" int ksfd_sweep(ksfd_ctx *c, const double *rin, double *x, double *rout, int first,
                          double norms[2], void *stream)
{
    TRY(check_ready(c));
    if (!c->have_jac) return fail("ksfd_jvp_setup has not been called");
    if (!rin || !x || !rout) return fail("ksfd_sweep: NULL vector");
    if (rin == rout || rin == x || x == rout) return fail("ksfd_sweep: the three vectors must be distinct");
    ksfd_ksp_opts o{};
    o.precond = 1;
    o.ksp_type = 1;
    if (!sweep_eligible(c, o))
        return fail("ksfd_sweep: needs the marching kernels (2-D/3-D, <= 4 ligands) and, on "
                    "several ranks, the peer-memory exchange");
    cudaStream_t st = (cudaStream_t)stream;
    TRY(gm_alloc(c));
    if (!c->sw_partial) CK(cudaMalloc(&c->sw_partial, sizeof(double) * 3 * KSFD_SWEEP_CTAS));
    GmStatus *hs = static_cast<GmStatus *>(c->gm_status);
    GmStatus *hsd = static_cast<GmStatus *>(c->gm_status_dev);
    hs->iters_done = hs->cycle_done = hs->final_ = 0;
    CK(cudaMemsetAsync(c->gmi, 0, sizeof(int) * GMI_INTS, st));
    c->pushed_vec = nullptr;
    const bool defer = c->p2p_on && c->nranks > 1 && tma_consumer(c);
    TRY(exchange(c, rin, c->dof, 1, st, nullptr, defer));
    const HostVec rh = make_hvec(c, rin, c->dof, 1);
    const HostVec ph = make_hvec(c, c->pc, 1, 2);
    const HostVec ch = coef_hvec(c);
    GmOpts go{0.0, 0.0, 0.0, 1 << 30, 1, 0, 0.0};
    SweepFin fin{c->sw_partial, KSFD_SWEEP_CTAS, 0, 1, 0, c->gm, c->gmi, hsd, go, 1e300, p2p_red(c),
                 c->gm_done};
    SweepHost a{x, rout, 1.0, first ? 1 : 0, KSFD_SWEEP_CTAS, &fin, nullptr};
    {
        ProfScope prof(c, 6, st);
        TRY(c->dim == 2 ? ksfd_march_sweep_d2(c, ch, rh, ph, a, nullptr, st)
                        : ksfd_march_sweep_d3(c, ch, rh, ph, a, nullptr, st));
    }
    if (norms) {
        CK(cudaStreamSynchronize(st));
        norms[0] = hs->rnorm0;
        norms[1] = hs->rnorm;
    }
    return 0;
}

static int gmres_impl(ksfd_ctx *c, const double *rhs, double rhs_sign, double *x,
                      const ksfd_ksp_opts &o, ksfd_ksp_result *res, cudaStream_t st)
{
    if (sweep_eligible(c, o)) {
        // automatic choice (ksp_type 2): after a fallback the next solves start with GMRES
        if (o.ksp_type == 2 && c->sw_backoff > 0)
            --c->sw_backoff;
        else
            return sweep_solve_impl(c, rhs, rhs_sign, x, o, res, st);
    }
    if (c->gm_pipeline && !o.reorth) return gmres_pipe_impl(c, rhs, rhs_sign, x, o, res, st);
    ksfd_ksp_opts o1 = o;                   // the host-driven variant knows block Jacobi only
    if (o1.precond > 1) o1.precond = 1;
    return gmres_sync_impl(c, rhs, rhs_sign, x, o1, res, st);
}

extern "C" int ksfd_gmres(ksfd_ctx *c, const double *rhs, double *x,
                          const ksfd_ksp_opts *o, ksfd_ksp_result *res, void *stream)
{
    TRY(check_ready(c));
    if (!c->have_jac) return fail("ksfd_jvp_setup has not been called");
    if (!rhs || !x || !o) return fail("ksfd_gmres: NULL argument");
    ksfd_ksp_opts o1 = *o;
    o1.ksp_type = 0;
    return gmres_impl(c, rhs, 1.0, x, o1, res, (cudaStream_t)stream);
}

extern "C" int ksfd_ksp_solve(ksfd_ctx *c, const double *rhs, double *x,
                              const ksfd_ksp_opts *o, ksfd_ksp_result *res, void *stream)
{
    TRY(check_ready(c));
    if (!c->have_jac) return fail("ksfd_jvp_setup has not been called");
    if (!rhs || !x || !o) return fail("ksfd_ksp_solve: NULL argument");
    if (o->ksp_type < 0 || o->ksp_type > 2) return fail("ksfd_ksp_solve: ksp_type must be 0, 1 or 2");
    if (o->ksp_type == 1 && !sweep_eligible(c, *o))
        return fail("ksfd_ksp_solve: richardson needs the marching kernels with the block-Jacobi "
                    "preconditioner (and peer-memory exchange on several ranks)");
    return gmres_impl(c, rhs, 1.0, x, *o, res, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------
// TS step.  ROSW 'ra34pw2' (Rang & Angermann 2005), transformed the way PETSc
// TSRosWRegister does: At = A*inv(Gamma), bt = b*inv(Gamma), stage solves
// with shift = 1/(h*gamma), Jacobian evaluated once at stage 0, -snes_type
// ksponly (one linear solve per stage, zero initial guess).
// ---------------------------------------------------------------------------
struct RoswTab {
    int s;
    double At[4][4], Gi[4][4], bt[4], bet[4], asum[4], gamma;
};

static const RoswTab &ra34pw2()
{
    static RoswTab T;
    static bool init = false;
    if (init) return T;
    const double g = 4.3586652150845900e-01;
    const double A[4][4] = {{0, 0, 0, 0},
                            {8.7173304301691801e-01, 0, 0, 0},
                            {8.4457060015369423e-01, -1.1299064236484185e-01, 0, 0},
                            {0, 0, 1.0, 0}};
    const double Gm[4][4] = {{g, 0, 0, 0},
                             {-8.7173304301691801e-01, g, 0, 0},
                             {-9.0338057013044082e-01, 5.4180672388095326e-02, g, 0},
                             {2.4212380706095346e-01, -1.2232505839045147e+00,
                              5.4526025533510214e-01, g}};
    const double b[4] = {2.4212380706095346e-01, -1.2232505839045147e+00,
                         1.5452602553351020e+00, g};
    const double be[4] = {3.7810903145819369e-01, -9.6042292212423178e-02,
                          5.0000000000000000e-01, 2.1793326075422950e-01};
    T.s = 4;
    T.gamma = g;
    // inverse of lower-triangular Gamma by forward substitution
    for (int col = 0; col < 4; ++col)
        for (int i = 0; i < 4; ++i) {
            double s = (i == col) ? 1.0 : 0.0;
            for (int k = 0; k < i; ++k) s -= Gm[i][k] * T.Gi[k][col];
            T.Gi[i][col] = s / Gm[i][i];
        }
    for (int i = 0; i < 4; ++i) {
        T.asum[i] = 0;
        for (int j = 0; j < 4; ++j) {
            T.asum[i] += A[i][j];
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[i][k] * T.Gi[k][j];
            T.At[i][j] = s;
        }
    }
    for (int j = 0; j < 4; ++j) {
        double s = 0, se = 0;
        for (int k = 0; k < 4; ++k) {
            s += b[k] * T.Gi[k][j];
            se += be[k] * T.Gi[k][j];
        }
        T.bt[j] = s;
        T.bet[j] = se;
    }
    init = true;
    return T;
}

template <int NV>
static int combine_launch(ksfd_ctx *c, const double *u, const VecList &Y, const CoefList &a,
                          const CoefList &gm, double *Z, double *Zd, cudaStream_t st)
{
    HaloPush hp{};
    if (c->fuse_push_mask & 2) hp = make_push(c, 0);
    if (hp.up_lo0) {
        c->pushed_vec0 = Z;
        k_stage_combine<NV, true><<<KSFD_RED_BLOCKS, 256, 0, st>>>(nlocal(c), u, Y, a, gm, Z, Zd, hp);
    } else {
        k_stage_combine<NV, false><<<KSFD_RED_BLOCKS, 256, 0, st>>>(nlocal(c), u, Y, a, gm, Z, Zd, hp);
    }
    CKL();
    return 0;
}

static int rosw_attempt(ksfd_ctx *c, const double *u, double t, double h,
                        const ksfd_ts_opts &o, const double *src, ksfd_time_cb cb,
                        void *user, double *unew, double *enorm, int *ksp_its,
                        int *ksp_fail, cudaStream_t st, bool velmax, double *accept_into)
{
    const RoswTab &T = ra34pw2();
    const long long n = nlocal(c);
    for (int i = 2; i <= 9; ++i) TRY(ensure_work(c, i));
    double *Y[4] = {c->work[2], c->work[3], c->work[4], c->work[5]};
    double *Z = c->work[6], *Zd = c->work[7];
    *ksp_fail = 0;
    // One rank, no stage callback: the work that FOLLOWS a stage solve — the next stage's
    // combination and residual, after the last stage the completion kernels — is enqueued
    // behind the predicted sweeps of the solve, before the host waits for its decision, so
    // the device does not idle during that round trip.  It is valid when the prediction held
    // (`clean`); otherwise it is simply done again.  The right-hand sides alternate between
    // two buffers: a solve that goes on (more sweeps, GMRES) still needs its own.
    const bool may_spec = (c->nranks == 1 || c->p2p_on) && !cb && c->spec_on;
    double *Fb[2] = {c->work[8], c->work[8]};
    if (may_spec) {
        TRY(ensure_work(c, 12));
        Fb[1] = c->work[12];
    }
    auto stage_pre = [&](int i, double *F, bool swp) -> int {
        const double ti = t + h * T.asum[i];
        VecList yl{};
        CoefList a{}, gm{};
        for (int j = 0; j < 4; ++j) yl.v[j] = Y[std::min(j, std::max(i - 1, 0))];
        for (int j = 0; j < i; ++j) {
            a.c[j] = T.At[i][j];
            gm.c[j] = T.Gi[i][j] / h;
        }
        const double *Zp = Z;
        switch (i) {
        case 0: Zp = u; CK(cudaMemsetAsync(Zd, 0, sizeof(double) * n, st)); break;
        case 1: TRY(combine_launch<1>(c, u, yl, a, gm, Z, Zd, st)); break;
        case 2: TRY(combine_launch<2>(c, u, yl, a, gm, Z, Zd, st)); break;
        default: TRY(combine_launch<3>(c, u, yl, a, gm, Z, Zd, st)); break;
        }
        if (cb) {
            CK(cudaStreamSynchronize(st));
            cb(ti, user);
        }
        return residual_impl(c, Zp, Zd, src, F, st, swp);   // F = Zdot - f(Z)
    };
    // completion, error norm and (velmax, KSFD_TS_VELOCITY_MAX) the CFL maxima of the step's
    // result, computed right away on the candidate (wasted only when the step is rejected):
    // they travel to the host with the error norm — one synchronisation per step
    auto step_tail = [&]() -> int {
        VecList yl{};
        CoefList b{}, be{};
        for (int j = 0; j < 4; ++j) {
            yl.v[j] = Y[j];
            b.c[j] = T.bt[j];
            be.c[j] = T.bet[j];
        }
        k_complete_step<4><<<KSFD_RED_BLOCKS, KSFD_RED_THREADS, 0, st>>>(
            n, u, yl, b, be, o.atol, o.rtol, unew, c->partial);
        CKL();
        k_reduce_partials<<<1, 128, 0, st>>>(1, KSFD_RED_BLOCKS, c->partial, c->dscal + SC_ENORM, 0);
        CKL();
        TRY(allreduce_dev(c, c->dscal + SC_ENORM, 1, ncclSum_, st));
        if (velmax) {
            TRY(velocity_impl(c, unew, nullptr, c->dscal + SC_ENORM + 1, st));
            TRY(allreduce_dev(c, c->dscal + SC_ENORM + 1, c->dim, ncclMax_, st));
        }
        return 0;
    };
    bool pre_done = false, tail_done = false;
    for (int i = 0; i < T.s; ++i) {
        double *F = Fb[i & 1];
        const bool swp = sweep_eligible(c, o.ksp) && !(o.ksp.ksp_type == 2 && c->sw_backoff > 0);
        if (!pre_done) TRY(stage_pre(i, F, swp));
        pre_done = false;
        if (i == 0) TRY(jvp_setup_impl(c, u, 1.0 / (h * T.gamma), nullptr, st, true));
        ksfd_ksp_result kr{};
        c->sw_slot = i;
        if (may_spec && swp) {
            SweepRun r;
            TRY(sweep_begin(c, r, F, -1.0, Y[i], o.ksp, st));       // A Y_i = -F
            const bool spec = sweep_settled(c, r);
            if (spec) {
                if (i + 1 < T.s)
                    TRY(stage_pre(i + 1, Fb[(i + 1) & 1], swp));
                else
                    TRY(step_tail());
                r.dirty = c->nranks > 1;
            }
            bool clean = false;
            TRY(sweep_end(c, r, &kr, &clean));
            if (spec && clean) {
                pre_done = i + 1 < T.s;
                tail_done = i + 1 == T.s;
            }
        } else {
            TRY(gmres_impl(c, F, -1.0, Y[i], o.ksp, &kr, st));      // A Y_i = -F
        }
        *ksp_its += kr.its;
        if (kr.reason < 0) {
            *ksp_fail = 1;
            return 0;
        }
    }
    if (!tail_done) TRY(step_tail());
    const int nfetch = 1 + (velmax ? c->dim : 0);
    // no step-size control: the candidate IS the new state — copied before the host waits
    if (accept_into)
        CK(cudaMemcpyAsync(accept_into, unew, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    TRY(fetch(c, SC_ENORM, nfetch, st));
    const double ntot = (double)c->g.plane_pts * (double)c->last_global * c->dof;
    *enorm = std::sqrt(c->hscal[SC_ENORM] / ntot);
    return 0;
}

static int beuler_attempt(ksfd_ctx *c, const double *u, double t, double h,
                          const ksfd_ts_opts &o, const double *src, ksfd_time_cb cb,
                          void *user, double *unew, int *ksp_its, int *ksp_fail,
                          cudaStream_t st)
{
    const long long n = nlocal(c);
    TRY(ensure_work(c, 2));
    TRY(ensure_work(c, 8));
    double *Y = c->work[2], *F = c->work[8];
    if (cb) {
        CK(cudaStreamSynchronize(st));
        cb(t + h, user);
    }
    TRY(residual_impl(c, u, nullptr, src, F, st));          // F = f(u)
    TRY(jvp_setup_impl(c, u, 1.0 / h, nullptr, st));
    ksfd_ksp_result kr{};
    TRY(gmres_impl(c, F, 1.0, Y, o.ksp, &kr, st));          // (I/h - J) Y = f(u)
    *ksp_its += kr.its;
    *ksp_fail = kr.reason < 0;
    CK(cudaMemcpyAsync(unew, u, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
    const double one = 1.0;
    const double *yp = Y;
    TRY(maxpy_impl(c, 1, &one, &yp, 1.0, unew, st));
    return 0;
}

extern "C" int ksfd_ts_step(ksfd_ctx *c, double *u, double t, double h,
                            const ksfd_ts_opts *o, const double *src, ksfd_time_cb cb,
                            void *user, ksfd_ts_result *res, void *stream)
{
    TRY(check_ready(c));
    if (!u || !o || !res) return fail("ksfd_ts_step: NULL argument");
    if (!(h > 0.0)) return fail("ksfd_ts_step: step size must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = nlocal(c);
    TRY(ensure_work(c, 9));
    double *unew = c->work[9];
    memset(res, 0, sizeof(*res));
    const bool velmax = (o->flags & KSFD_TS_VELOCITY_MAX) != 0;
    if (o->flags & KSFD_TS_GROOM) {
        k_groom<<<nblk(n, 256), 256, 0, st>>>(c->g, c->P.rhomin, c->P.Umin, u);
        CKL();
    }
    const double safety0 = o->safety > 0 ? o->safety : 0.9;
    const double rsafety = o->reject_safety > 0 ? o->reject_safety : 0.5;
    const double clip_lo = o->clip_lo > 0 ? o->clip_lo : 0.1;
    const double clip_hi = o->clip_hi > 0 ? o->clip_hi : 10.0;
    const double dt_min = o->dt_min > 0 ? o->dt_min : 1e-20;
    const double dt_max = o->dt_max > 0 ? o->dt_max : 1e50;
    const int max_rej = o->max_reject > 0 ? o->max_reject : 10;
    const int order = 3;
    while (true) {
        double enorm = 0.0;
        int fail_ = 0;
        if (o->ts_type == 1)
            TRY(beuler_attempt(c, u, t, h, *o, src, cb, user, unew, &res->ksp_its, &fail_, st));
        else
            TRY(rosw_attempt(c, u, t, h, *o, src, cb, user, unew, &enorm,
                             &res->ksp_its, &fail_, st, velmax, o->adapt == 1 ? nullptr : u));
        if (fail_) {
            res->ksp_fail = 1;
            res->accepted = 0;
            res->h_used = h;
            res->h_next = h;
            res->t_new = t;
            return 0;
        }
        res->enorm = enorm;
        bool accept = true;
        double hnext = h;
        if (o->adapt == 1 && o->ts_type != 1) {
            // TSAdaptChoose_Basic
            double safety = safety0;
            if (enorm > 1.0) {
                accept = false;
                safety *= rsafety;
            }
            double hfac = enorm > 0.0 ? safety * std::pow(enorm, -1.0 / order) : clip_hi;
            hfac = std::min(std::max(hfac, clip_lo), clip_hi);
            hnext = std::min(std::max(h * hfac, dt_min), dt_max);
        }
        if (accept) {
            if (o->adapt == 1 || o->ts_type == 1)
                CK(cudaMemcpyAsync(u, unew, sizeof(double) * n, cudaMemcpyDeviceToDevice, st));
            if (velmax) {
                if (o->ts_type == 1) {      // backward Euler has no error-norm read to ride on
                    TRY(velocity_impl(c, unew, nullptr, c->dscal + SC_ENORM + 1, st));
                    TRY(allreduce_dev(c, c->dscal + SC_ENORM + 1, c->dim, ncclMax_, st));
                    TRY(fetch(c, SC_ENORM + 1, c->dim, st));
                }
                for (int d = 0; d < c->dim; ++d) res->vmax[d] = c->hscal[SC_ENORM + 1 + d];
                res->have_vmax = 1;
            }
            res->accepted = 1;
            res->t_new = t + h;
            res->h_used = h;
            res->h_next = hnext;
            return 0;
        }
        ++res->rejections;
        if (res->rejections > max_rej || hnext < dt_min) {
            res->accepted = 0;
            res->t_new = t;
            res->h_used = h;
            res->h_next = hnext;
            return 0;
        }
        h = hnext;
    }
}
