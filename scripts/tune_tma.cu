// Stand-alone tuner / checker for the TMA-fed marching kernels (tma_march.cuh) against the
// register-prefetch kernels (march_kernels.cuh).  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo
//        -I ksfd_b200/csrc scripts/tune_tma.cu -o gpurun_out/tune_tma
//   ./tune_tma [2d N | 3d N] ...
// Every TMA variant is compared bit for bit with the reference kernel of the same operator.
// TUNE_ONLY="i,j": run only the variants with these ordinals (for ncu).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tma_march.cuh"
#include "tma_host.h"

#define CHECK(x)                                                                    \
    do {                                                                            \
        cudaError_t e_ = (x);                                                       \
        if (e_ != cudaSuccess) {                                                    \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(1);                                                                \
        }                                                                           \
    } while (0)

static DevPhys make_phys(int dim)
{
    DevPhys P{};
    P.ngroups = 2;
    P.nlig = 2;
    P.cap_type = 0;
    P.dim = dim;
    P.s2 = 0.02357 * 0.02357 / 2;
    P.rhomax = 28000.0;
    P.inv_cushion = 1.0 / 2000.0;
    P.capscale = 2.0 * P.s2;
    P.rhomin = 1e-7;
    P.Umin = 1e-7;
    P.inv_rhomax = 1.0 / 28000.0;
    P.alpha[0] = P.alpha[1] = 1500.0;
    P.beta[0] = 5.56e-4;
    P.beta[1] = -5.56e-4;
    P.lig_group[0] = 0;
    P.lig_group[1] = 1;
    P.weight[0] = P.weight[1] = 1.0;
    P.Wgl[0][0] = 1.0;
    P.Wgl[1][1] = 1.0;
    P.s[0] = P.gamma[0] = 0.01;
    P.D[0] = 1e-6;
    P.s[1] = P.gamma[1] = 0.001;
    P.D[1] = 1e-5;
    const double h = 1.0 / 384;
    const double a1[5] = {1, -8, 0, 8, -1}, a2[5] = {-1, 16, -30, 16, -1};
    for (int a = 0; a < dim; ++a)
        for (int s = 0; s < 5; ++s) {
            P.w1[a][s] = a1[s] / (12 * h);
            P.w2[a][s] = a2[s] / (12 * h * h);
        }
    P.w2c = dim * a2[2] / (12 * h * h);
    for (int a = 0; a < dim; ++a) {
        P.c1[a] = 1 / (12 * h);
        P.c2[a] = 1 / (12 * h * h);
        P.c1sq[a] = P.c1[a] * P.c1[a];
    }
    P.sym_ok = 1;
    P.ycap1 = -2.0 * P.inv_cushion;
    P.ycap0 = 2.0 * P.rhomax * P.inv_cushion;
    P.capscale2 = 2.0 * P.capscale;
    P.mk = fastk_default();
    return P;
}

__global__ void k_fill(double *p, long long n, double base, double amp, unsigned seed)
{
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long x = (i + 1) * 0x9E3779B97F4A7C15ull + seed * 0xD1B54A32D192ED03ull;
    x ^= x >> 29;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 32;
    p[i] = base + amp * ((double)(x & 0xFFFFFF) / 16777216.0 - 0.5);
}

// number of elements whose bit patterns differ, and the largest difference
__global__ void k_diff(const double *a, const double *b, long long n, unsigned long long *cnt,
                       double *maxd)
{
    unsigned long long c = 0;
    double m = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        if (__double_as_longlong(a[i]) != __double_as_longlong(b[i])) {
            ++c;
            double d = fabs(a[i] - b[i]);
            if (!(d <= m)) m = d;       // NaN propagates
        }
    }
    if (c) {
        atomicAdd(cnt, c);
        atomicMax(reinterpret_cast<unsigned long long *>(maxd),
                  (unsigned long long)__double_as_longlong(m));
    }
}

struct Problem {
    int dim, n0, n1, nloc;
    long long npts;
    int nrot, nout;
    std::vector<double *> u, v, out;
    double *coef, *pc, *ref;
};

static int g_ordinal = 0;
static bool selected()
{
    const int me = g_ordinal++;
    const char *f = getenv("TUNE_ONLY");
    if (!f || !*f) return true;
    char buf[256];
    strncpy(buf, f, 255);
    buf[255] = 0;
    for (char *t = strtok(buf, ","); t; t = strtok(nullptr, ","))
        if (atoi(t) == me) return true;
    return false;
}

static int pick_rz(int rz, long long cols, int occ, int nloc)
{
    if (rz <= 0) {
        // auto: fill occ*148 slots in (-rz or 1) whole waves
        const int slots = 148 * occ;
        int waves = (int)((cols + slots - 1) / slots);
        int nch = (int)((long long)waves * slots / cols);
        if (nch < 1) nch = 1;
        if (-rz > 1) nch *= -rz;
        rz = (nloc + nch - 1) / nch;
        if (rz < 2) rz = 2;
    }
    if (rz > nloc) rz = nloc;
    return rz;
}

template <class F>
static double time_launches(F launch, int nrot)
{
    cudaEvent_t e0, e1;
    CHECK(cudaEventCreate(&e0));
    CHECK(cudaEventCreate(&e1));
    const int reps = getenv("TUNE_ONLY") ? 2 : 20;
    for (int i = 0; i < 3; ++i) launch(i % nrot);
    CHECK(cudaDeviceSynchronize());
    CHECK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch(i % nrot);
    CHECK(cudaEventRecord(e1));
    CHECK(cudaDeviceSynchronize());
    float ms = 0;
    CHECK(cudaEventElapsedTime(&ms, e0, e1));
    CHECK(cudaEventDestroy(e0));
    CHECK(cudaEventDestroy(e1));
    return ms * 1e3 / reps;
}

// reference kernels: record out[0] of one launch on buffer set 0 in pb.ref
template <int DIM, int TX, int TY, class Op, int MINB, bool UNR, int RINGS = 2>
static void run_ref(const char *name, Problem &pb, const DevPhys &P, Op op_proto,
                    void (*bind)(Op &, const Problem &, int), bool record, int waves = 0)
{
    using T = TileT<DIM, TX, TY>;
    const int ord = g_ordinal;
    if (!selected() && !record) return;
    auto kern = k_march<DIM, TX, TY, Op, MINB, UNR, RINGS>;
    const size_t smem = march_smem_bytes<Op, T::SP, RINGS>();
    CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T::NT, smem));
    cudaFuncAttributes fa;
    CHECK(cudaFuncGetAttributes(&fa, kern));
    const int ntx = (pb.n0 + TX - 1) / TX, ox = (pb.n0 + ntx - 1) / ntx;
    const int nty = DIM == 3 ? (pb.n1 + TY - 1) / TY : 1, oy = DIM == 3 ? (pb.n1 + nty - 1) / nty : 1;
    if (occ < 1) {
        printf("#%02d %-12s REF TX%3d TY%2d MINB%d R%d: occupancy 0\n", ord, name, TX, TY, MINB, RINGS);
        return;
    }
    const int rz = pick_rz(-waves, (long long)ntx * nty, occ, pb.nloc);
    const int nch = (pb.nloc + rz - 1) / rz;
    MarchArgs a{pb.n0, pb.n1, pb.nloc, pb.n0 * pb.n1, ox, oy, rz};
    dim3 grid(ntx, nty, nch);
    auto launch = [&](int i) {
        Op op = op_proto;
        bind(op, pb, i);
        kern<<<grid, T::NT, smem>>>(a, P, op, nullptr);
    };
    const double us = time_launches(launch, pb.nrot);
    unsigned long long hcnt = 0;
    if (record) {
        launch(0);
        CHECK(cudaMemcpy(pb.ref, pb.out[0], pb.npts * pb.nout * 8, cudaMemcpyDeviceToDevice));
    } else {
        CHECK(cudaMemset(pb.out[0], 0xff, pb.npts * pb.nout * 8));
        launch(0);
        unsigned long long *dcnt;
        double *dmax;
        CHECK(cudaMalloc(&dcnt, 8));
        CHECK(cudaMalloc(&dmax, 8));
        CHECK(cudaMemset(dcnt, 0, 8));
        CHECK(cudaMemset(dmax, 0, 8));
        k_diff<<<592, 256>>>(pb.out[0], pb.ref, pb.npts * pb.nout, dcnt, dmax);
        CHECK(cudaMemcpy(&hcnt, dcnt, 8, cudaMemcpyDeviceToHost));
        CHECK(cudaFree(dcnt));
        CHECK(cudaFree(dmax));
    }
    CHECK(cudaDeviceSynchronize());
    const double gpts = pb.npts / us / 1e3;
    printf("#%02d %-12s REF TX%3d TY%2d MINB%d UNR%d R%d regs%3d occ%d rz%4d grid %4dx%3dx%4d smem %6zu  %9.2f us  %6.2f Gpts/s  frac %.3f  diff %llu %s\n",
           ord, name, TX, TY, MINB, (int)UNR, RINGS, fa.numRegs, occ, rz, ntx, nty, nch, smem, us, gpts,
           gpts * 72.0 / 6544.7, hcnt, record ? "(recorded)" : hcnt ? "MISMATCH" : "ok");
    fflush(stdout);
}

// TMA kernels
template <int DIM, int TX, int TY, class Op, int MINB, bool UNR, int SC, int SH>
static void run_tma(const char *name, Problem &pb, const DevPhys &P, Op op_proto,
                    void (*bind)(Op &, const Problem &, int),
                    void (*bind_tma)(TmaInT<Op::NIN> &, const Problem &, int, int, int),
                    const int *rzs, int nrz)
{
    using M = TmaMarcher<DIM, TX, TY, Op, UNR, SC, SH>;
    const int ord = g_ordinal;
    if (!selected()) return;
    if (getenv("TUNE_ONLY")) nrz = 1;
    auto kern = k_tma_march<DIM, TX, TY, Op, MINB, UNR, SC, SH>;
    const size_t smem = tma_march_smem_bytes<DIM, TX, TY, Op, UNR, SC, SH>();
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
        cudaGetLastError();
        printf("#%02d %-12s TMA TX%3d TY%2d SC%d SH%d: %zu bytes of shared memory do not fit\n", ord,
               name, TX, TY, SC, SH, smem);
        return;
    }
    int occ = 0;
    CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, M::NTH, smem));
    cudaFuncAttributes fa;
    CHECK(cudaFuncGetAttributes(&fa, kern));
    if (occ < 1) {
        printf("#%02d %-12s TMA TX%3d TY%2d: occupancy 0\n", ord, name, TX, TY);
        return;
    }
    const int ntx = (pb.n0 + TX - 1) / TX;
    int ox = (pb.n0 + ntx - 1) / ntx;
    ox += ox & 1;
    const int nty = DIM == 3 ? (pb.n1 + TY - 1) / TY : 1;
    int oy = DIM == 3 ? (pb.n1 + nty - 1) / nty : 1;
    if (DIM == 3) oy += oy & 1;
    // tensor maps per buffer set (host encode + upload once)
    std::vector<TmaInT<Op::NIN>> tins(pb.nrot);
    for (int i = 0; i < pb.nrot; ++i) bind_tma(tins[i], pb, i, TX, TY);
    for (int r = 0; r < nrz; ++r) {
        const int rz = pick_rz(rzs[r], (long long)ntx * nty, occ, pb.nloc);
        const int nch = (pb.nloc + rz - 1) / rz;
        MarchArgs a{pb.n0, pb.n1, pb.nloc, pb.n0 * pb.n1, ox, oy, rz};
        dim3 grid(ntx, nty, nch);
        auto launch = [&](int i) {
            Op op = op_proto;
            bind(op, pb, i);
            kern<<<grid, M::NTH, smem>>>(a, P, op, tins[i], nullptr);
        };
        const double us = time_launches(launch, pb.nrot);
        CHECK(cudaGetLastError());
        // compare with the reference output
        CHECK(cudaMemset(pb.out[0], 0xff, pb.npts * pb.nout * 8));
        launch(0);
        unsigned long long *dcnt, hcnt = 0;
        double *dmax, hmax = 0;
        CHECK(cudaMalloc(&dcnt, 8));
        CHECK(cudaMalloc(&dmax, 8));
        CHECK(cudaMemset(dcnt, 0, 8));
        CHECK(cudaMemset(dmax, 0, 8));
        k_diff<<<592, 256>>>(pb.out[0], pb.ref, pb.npts * pb.nout, dcnt, dmax);
        CHECK(cudaMemcpy(&hcnt, dcnt, 8, cudaMemcpyDeviceToHost));
        CHECK(cudaMemcpy(&hmax, dmax, 8, cudaMemcpyDeviceToHost));
        CHECK(cudaFree(dcnt));
        CHECK(cudaFree(dmax));
        const double gpts = pb.npts / us / 1e3;
        printf("#%02d %-12s TMA TX%3d TY%2d MINB%d UNR%d SC%d SH%d regs%3d occ%d rz%4d grid %4dx%3dx%4d smem %6zu  %9.2f us  %6.2f Gpts/s  frac %.3f  diff %llu max %.3e %s\n",
               ord, name, TX, TY, MINB, (int)UNR, SC, SH, fa.numRegs, occ, rz, ntx, nty, nch, smem,
               us, gpts, gpts * 72.0 / 6544.7, hcnt, hmax, hcnt ? "MISMATCH" : "ok");
        fflush(stdout);
    }
}

template <int DIM>
static void bind_res(ResidualOp<DIM, 2, true> &op, const Problem &pb, int i)
{
    op.u.base = pb.u[i];
    op.u.lo = pb.u[i] + (long long)(pb.nloc - 2) * pb.n0 * pb.n1 * 3;
    op.u.hi = pb.u[i];
    op.u.par = nullptr;
    op.u.pstride = 0;
    op.udot = pb.v[i];
    op.src = nullptr;
    op.out = pb.out[i];
}
template <int DIM, bool PC>
static void bind_jvp(JvpOp<DIM, 2, PC> &op, const Problem &pb, int i)
{
    const long long ps = (long long)pb.n0 * pb.n1;
    op.coef.lo = pb.coef;
    op.coef.base = pb.coef + 2 * ps * 5;
    op.coef.hi = pb.coef + (2 + pb.nloc) * ps * 5;
    op.coef.par = op.v.par = op.pc.par = nullptr;
    op.coef.pstride = op.v.pstride = op.pc.pstride = 0;
    op.v.base = pb.v[i];
    op.v.lo = pb.v[i] + (long long)(pb.nloc - 2) * ps * 3;
    op.v.hi = pb.v[i];
    op.pc.base = pb.pc;
    op.pc.lo = pb.pc + (long long)(pb.nloc - 2) * ps;
    op.pc.hi = pb.pc;
    op.shift = 2294.0;
    op.invd[0] = 1.0 / 2294.5;
    op.invd[1] = 1.0 / 2294.7;
    op.out = pb.out[i];
}

static CUtensorMap *g_dmaps = nullptr;      // device pool of tensor maps (global-memory variant)
static int g_dmaps_used = 0;
// encode the three box shapes of one buffer; returns where the kernel finds them
static void put_maps(void *dst, const double *base, const Problem &pb, long long nfp, int nc,
                     int TX, int TY)
{
    CUtensorMap h[3];
    std::string e = ksfd_make_tmaps(h, base, pb.n0, pb.n1, nfp, nc, TX, TY);
    if (!e.empty()) {
        printf("%s\n", e.c_str());
        exit(1);
    }
#if KSFD_TMAP_PARAM
    memcpy(dst, h, sizeof(h));
#else
    if (!g_dmaps) CHECK(cudaMalloc(&g_dmaps, sizeof(CUtensorMap) * 3 * 4096));
    if (g_dmaps_used + 3 > 3 * 4096) g_dmaps_used = 0;
    CUtensorMap *d = g_dmaps + g_dmaps_used;
    g_dmaps_used += 3;
    CHECK(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
    *reinterpret_cast<const CUtensorMap **>(dst) = d;
#endif
}
#if KSFD_TMAP_PARAM
#define MAPSLOT(t, i, set) ((void *)(t).m[i][set])
#define MAPCOPY(t, i) memcpy((t).m[i][1], (t).m[i][0], sizeof((t).m[i][0]))
#else
#define MAPSLOT(t, i, set) ((void *)&(t).m[i][set])
#define MAPCOPY(t, i) ((t).m[i][1] = (t).m[i][0])
#endif
static void tma_res(TmaInT<1> &t, const Problem &pb, int i, int TX, int TY)
{
    memset(&t, 0, sizeof(t));
    put_maps(MAPSLOT(t, 0, 0), pb.u[i], pb, 3LL * pb.nloc, 3, TX, TY);
    MAPCOPY(t, 0);
    t.v[0].wrap = 1;
    t.v[0].nc = 3;
    t.v[0].coff = 0;
}
template <int NIN>
static void tma_jvp(TmaInT<NIN> &t, const Problem &pb, int i, int TX, int TY)
{
    memset(&t, 0, sizeof(t));
    // coef is stored ghosted: plane 0 is plane 2 of the buffer, its ghost planes are in place
    put_maps(MAPSLOT(t, 0, 0), pb.coef, pb, 5LL * (pb.nloc + 4), 5, TX, TY);
    MAPCOPY(t, 0);
    t.v[0].kofs[0] = 2 * 5;
    t.v[0].kofs[1] = 0;
    t.v[0].kofs[2] = (2 + pb.nloc) * 5;
    t.v[0].wrap = 0;
    t.v[0].nc = 5;
    t.v[0].coff = 0;
    put_maps(MAPSLOT(t, 1, 0), pb.v[i], pb, 3LL * pb.nloc, 3, TX, TY);
    MAPCOPY(t, 1);
    t.v[1].wrap = 1;
    t.v[1].nc = 3;
    t.v[1].coff = 5;
    if (NIN > 2) {
        put_maps(MAPSLOT(t, NIN - 1, 0), pb.pc, pb, 1LL * pb.nloc, 1, TX, TY);
        MAPCOPY(t, NIN - 1);
        t.v[NIN - 1].wrap = 1;
        t.v[NIN - 1].nc = 1;
        t.v[NIN - 1].coff = 8;
    }
}

static Problem make_problem(int dim, int n)
{
    Problem pb;
    pb.dim = dim;
    pb.n0 = n;
    pb.n1 = dim == 3 ? n : 1;
    pb.nloc = n;
    pb.nout = 3;
    pb.npts = (long long)pb.n0 * pb.n1 * pb.nloc;
    const long long N = pb.npts * 3;
    pb.nrot = (int)(2.5 * 126e6 / (N * 8 * 3)) + 1;
    if (pb.nrot < 2) pb.nrot = 2;
    if (pb.nrot > 12) pb.nrot = 12;
    for (int i = 0; i < pb.nrot; ++i) {
        double *u, *v, *o;
        CHECK(cudaMalloc(&u, N * 8));
        CHECK(cudaMalloc(&v, N * 8));
        CHECK(cudaMalloc(&o, N * 8));
        k_fill<<<(unsigned)((N + 255) / 256), 256>>>(u, N, 9000.0, 300.0, 3 * i);
        k_fill<<<(unsigned)((N + 255) / 256), 256>>>(v, N, 0.0, 2.0, 3 * i + 1);
        pb.u.push_back(u);
        pb.v.push_back(v);
        pb.out.push_back(o);
    }
    const long long NC = (long long)(pb.nloc + 4) * pb.n0 * pb.n1 * 5;
    CHECK(cudaMalloc(&pb.coef, NC * 8));
    k_fill<<<(unsigned)((NC + 255) / 256), 256>>>(pb.coef, NC, 1.0, 0.5, 77);
    CHECK(cudaMalloc(&pb.pc, pb.npts * 8));
    k_fill<<<(unsigned)((pb.npts + 255) / 256), 256>>>(pb.pc, pb.npts, 4e-4, 1e-5, 78);
    CHECK(cudaMalloc(&pb.ref, N * 8));
    CHECK(cudaDeviceSynchronize());
    return pb;
}

static void free_problem(Problem &pb)
{
    for (auto p : pb.u) cudaFree(p);
    for (auto p : pb.v) cudaFree(p);
    for (auto p : pb.out) cudaFree(p);
    cudaFree(pb.coef);
    cudaFree(pb.pc);
    cudaFree(pb.ref);
}

#define REF_RES(DIM, TX, TY, MINB, REC) \
    run_ref<DIM, TX, TY, ResidualOp<DIM, 2, true>, MINB, false>("residual", pb, P, ResidualOp<DIM, 2, true>{}, bind_res<DIM>, REC)
#define REF_JVP(DIM, TX, TY, MINB, PC, REC) \
    run_ref<DIM, TX, TY, JvpOp<DIM, 2, PC>, MINB, true>(PC ? "jvp_pc" : "jvp", pb, P, JvpOp<DIM, 2, PC>{}, bind_jvp<DIM, PC>, REC)
// R4: four ring slots, the stencil lags the shared plane by two barriers
#define R4_RES(DIM, TX, TY, MINB, UNR, W) \
    run_ref<DIM, TX, TY, ResidualOp<DIM, 2, true>, MINB, UNR, 4>("residual", pb, P, ResidualOp<DIM, 2, true>{}, bind_res<DIM>, false, W)
#define R4_JVP(DIM, TX, TY, MINB, UNR, PC, W) \
    run_ref<DIM, TX, TY, JvpOp<DIM, 2, PC>, MINB, UNR, 4>(PC ? "jvp_pc" : "jvp", pb, P, JvpOp<DIM, 2, PC>{}, bind_jvp<DIM, PC>, false, W)
#define TMA_RES(DIM, TX, TY, MINB, UNR, SC, SH) \
    run_tma<DIM, TX, TY, ResidualOp<DIM, 2, true>, MINB, UNR, SC, SH>("residual", pb, P, ResidualOp<DIM, 2, true>{}, bind_res<DIM>, tma_res, rzs, nrz)
#define TMA_JVP(DIM, TX, TY, MINB, UNR, PC, SC, SH) \
    run_tma<DIM, TX, TY, JvpOp<DIM, 2, PC>, MINB, UNR, SC, SH>(PC ? "jvp_pc" : "jvp", pb, P, JvpOp<DIM, 2, PC>{}, bind_jvp<DIM, PC>, tma_jvp<PC ? 3 : 2>, rzs, nrz)

int main(int argc, char **argv)
{
    const int rzs[] = {0, -2, -3};     // 1, 2, 3 waves of CTAs
    const int nrz = 3;
    for (int a = 1; a + 1 < argc; a += 2) {
        const int dim = argv[a][0] == '3' ? 3 : 2;
        const int n = atoi(argv[a + 1]);
        Problem pb = make_problem(dim, n);
        DevPhys P = make_phys(dim);
        printf("== %dD n=%d  (%lld points, %d buffer sets)\n", dim, n, pb.npts, pb.nrot);
        const bool more = getenv("TUNE_MORE") != nullptr;       // only the extra variants
        if (dim == 2 && !more) {
            REF_RES(2, 124, 1, 6, true);
            TMA_RES(2, 128, 1, 6, false, 3, 3);
            TMA_RES(2, 256, 1, 3, false, 3, 3);
            REF_JVP(2, 124, 1, 4, true, true);
            TMA_JVP(2, 128, 1, 4, true, true, 3, 3);
            TMA_JVP(2, 128, 1, 4, true, true, 2, 2);
            TMA_JVP(2, 256, 1, 2, true, true, 3, 3);
            REF_JVP(2, 124, 1, 4, false, true);
            TMA_JVP(2, 128, 1, 4, true, false, 3, 3);
            TMA_JVP(2, 256, 1, 2, true, false, 3, 3);
        } else if (dim == 2) {
            {
                REF_JVP(2, 124, 1, 4, true, true);
                TMA_JVP(2, 256, 1, 2, true, true, 3, 3);
                TMA_JVP(2, 256, 1, 2, false, true, 3, 3);
                TMA_JVP(2, 256, 1, 2, true, true, 2, 2);
                TMA_JVP(2, 256, 1, 2, true, true, 4, 4);
                TMA_JVP(2, 128, 1, 4, false, true, 3, 3);
            }
        } else if (!more) {
            REF_RES(3, 16, 16, 2, true);
            TMA_RES(3, 16, 16, 2, false, 3, 3);
            TMA_RES(3, 16, 8, 4, false, 3, 3);
            REF_JVP(3, 16, 16, 1, true, true);
            TMA_JVP(3, 16, 16, 2, true, true, 2, 2);
            TMA_JVP(3, 16, 16, 2, true, true, 3, 2);
            TMA_JVP(3, 16, 8, 4, true, true, 2, 2);
            TMA_JVP(3, 16, 8, 3, true, true, 3, 3);
            TMA_JVP(3, 32, 8, 2, true, true, 2, 2);
            REF_JVP(3, 16, 16, 1, false, true);
            TMA_JVP(3, 16, 16, 2, true, false, 2, 2);
            TMA_JVP(3, 16, 16, 2, true, false, 3, 3);
            TMA_JVP(3, 16, 8, 4, true, false, 2, 2);
        } else {
            {
                REF_JVP(3, 16, 16, 1, true, true);
                TMA_JVP(3, 16, 16, 2, true, true, 2, 2);
                TMA_JVP(3, 16, 16, 2, false, true, 2, 2);
                TMA_JVP(3, 16, 16, 2, true, true, 2, 3);
                TMA_JVP(3, 32, 16, 1, true, true, 2, 2);
                TMA_JVP(3, 16, 32, 1, true, true, 2, 2);
                TMA_JVP(3, 8, 32, 2, true, true, 2, 2);
            }
        }
        free_problem(pb);
    }
    return 0;
}
