// Host-side context shared by the translation units of libksfd_b200.so.
#pragma once
#include <cstdint>
#include <string>
#include <cuda_runtime.h>

#include "device_common.cuh"

// ---------------------------------------------------------------------------
// errors (thread-local last message; every entry point returns 0 / non-zero)
// ---------------------------------------------------------------------------
int ksfd_fail(const std::string &m);
void ksfd_count_launch();
#define fail ksfd_fail
#define CK(call)                                                              \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess)                                                \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
#define CKL()                                                                 \
    do {                                                                      \
        ksfd_count_launch();                                                  \
        cudaError_t e_ = cudaGetLastError();                                  \
        if (e_ != cudaSuccess)                                                \
            return fail(std::string("kernel launch: ") +                      \
                        cudaGetErrorString(e_) + " at " + __FILE__ + ":" +    \
                        std::to_string(__LINE__));                            \
    } while (0)
#define TRY(x)                 \
    do {                       \
        int r_ = (x);          \
        if (r_) return r_;     \
    } while (0)

typedef struct ncclComm *ncclComm_t;

#define KSFD_HALO_SLOTS 4
#define KSFD_NSCAL 512          // device/host scalar scratch

#define KSFD_KLAUNCH(k, g, b, sm, st, ...) k<<<g, b, sm, st>>>(__VA_ARGS__)

struct ksfd_ctx {
    int dim = 0, dof = 0, device = 0;
    long long n[3] = {1, 1, 1};
    long long last_start = 0, last_count = 0, last_global = 0;
    Geom g{};
    DevPhys P{};
    // physics of the current LINEARISATION: snapshot taken by ksfd_jvp_setup and used by
    // every Jacobian-side kernel (J.v, preconditioners, solution update), so that
    // time-dependent parameters refreshed at the later ROSW stages (stage callback ->
    // ksfd_set_physics) change the residual only — PETSc's ROSW keeps the stage-0
    // Jacobian for the whole step
    DevPhys Pjac{};
    bool have_phys = false;
    // options
    int variant = 0, opt_tx = -1, opt_rz = 0;
    // in-situ kernel timing (option "profile"): event pairs around the stencil launches
    bool prof_on = false;
    void *prof = nullptr;        // ProfState* (ksfd.cu)
    bool opt_tile_set = false;
    // comm
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
    // halo slots: [lo(2 planes) | hi(2 planes)] per slot, sized for dof+2 stride
    double *halo[KSFD_HALO_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    size_t halo_plane_doubles = 0;
    // direct peer-to-peer halo push over NVLink (ksfd_p2p_export/import): one
    // IPC-shared allocation per rank: [64 doubles of flags][slot][parity][lo|hi]
    double *p2p_mine = nullptr, *p2p_dn = nullptr, *p2p_up = nullptr;
    double *p2p_peer[16] = {nullptr};       // every rank's allocation (own included), <= 16 ranks
    bool p2p_on = false;
    // device-side exchange counters: [0..3] halo slots, [4] all-reduce
    unsigned long long *p2p_ctr = nullptr;
    unsigned *p2p_done = nullptr;           // [0] exchange kernels, [1..2] fused producer pushes (combined / bottom)
    // Krylov vector whose boundary planes were pushed by the kernel that produced it
    // (consumed by the next jvp_impl on that vector; host-side bookkeeping only)
    const double *pushed_vec = nullptr;
    const double *pushed_vec0 = nullptr;    // same for halo slot 0 (stage vector Z -> residual)
    bool gm_no_push = false;
    int *p2p_err = nullptr, *p2p_err_dev = nullptr;    // pinned + mapped: a peer wait timed out
    // Jacobian state
    double *coef = nullptr;      // ghosted (nloc+4 planes) x (dof+2)
    double *pc = nullptr;        // nloc planes x 1: inverse Schur pivot per point
    double shift = 0.0;
    double invd[KSFD_MAX_LIGANDS] = {0};   // 1/(shift + gamma_l - D_l*w2c)
    bool have_jac = false;
    // reductions
    double *partial = nullptr;   // [KSFD_MAXV+1][KSFD_RED_BLOCKS]
    double *dscal = nullptr;     // device scalars (KSFD_NSCAL)
    double *hscal = nullptr;     // pinned host scalars (KSFD_NSCAL)
    void *plan_cache = nullptr;  // std::map<long long, MarchPlan>*
    void *tmap_cache = nullptr;  // TmapCache* (march_launch.cuh): encoded tensor maps
    // pipelined GMRES: device state, pinned host-visible status
    double *gm = nullptr;
    int *gmi = nullptr;
    void *gm_status = nullptr, *gm_status_dev = nullptr;   // GmStatus (mapped)
    unsigned *gm_done = nullptr;                            // block counter of the fused multi-dot
    int gm_pipeline = 1, gm_runahead = 2;
    int gm_pred[2] = {0, 0};     // columns of the last first / later cycle (launch-ahead hint)
    // Richardson sweeps (sweep_op.cuh): per-CTA partial sums, launch-ahead hint, and the
    // automatic choice: after a solve that had to fall back to GMRES the next
    // `sw_backoff` solves start with GMRES directly
    double *sw_partial = nullptr;
    // sweeps of the last solve in each SLOT: the time integrator numbers its solves within a
    // step (ROSW: stage 0..3) and stage i of one step predicts stage i of the next
    int sw_hist[4] = {0, 0, 0, 0}, sw_slot = 0, sw_backoff = 0;
    int sw_stable[4] = {0, 0, 0, 0};   // consecutive solves of the slot with the predicted count
    double sw_slow = 0.35;
    // Every sweep from `lead` sweeps before the predicted end on is tested for convergence:
    // normally only the predicted last one (a test costs a serial epilogue of 7-12 us),
    // every 8th solve of a slot one more, so that the prediction can also go DOWN
    int sw_test_lead = 1;
    // one rank: the next stage's combination and residual (after the last stage: the completion
    // kernels) are enqueued behind the predicted sweeps of a solve, before the host waits for it
    bool spec_on = true;
    int sw_underpredict = 0;     // test knob: sweeps launched fewer than predicted
    int fuse_push_mask = 7;      // producers that push their output's boundary planes: 2 stage combination, 4 stage residual
    bool sw_fuse_push = true;    // several ranks: boundary planes pushed by the sweep kernel itself
    // Single-pass classical Gram-Schmidt loses orthogonality like eps*kappa^2,
    // kappa ~ the residual reduction inside the cycle, so a cycle is closed
    // after this reduction and restarted from the TRUE residual (measured on
    // the benchmark problem: with 1e-9 some stage solves need 20-80 iterations
    // instead of 6)
    double gm_cycle_factor = 1e-5;
    // spectral (FFT) preconditioner: cuFFT plans, spectra, coefficient means
    int fft_fwd = -1, fft_inv = -1;
    void *fft_spec = nullptr;    // double2 [dof][n2][n1][n0/2+1]
    double *fft_means = nullptr;
    // slab-distributed variant (several ranks, KSFD_FFT_MULTI=1): plane transforms,
    // last-axis transform, two spectrum buffers (fft_spec, fft_spec2)
    int fft_z = -1;
    void *fft_spec2 = nullptr;
    bool fft_dist = false;
    long long fft_ps = 0;        // plane wave numbers (n0/2+1, times n1 in 3-D)
    bool fft_failed = false;
    bool fft_means_valid = false;   // means belong to the current linearisation
    bool pc_auto_fft = false;    // precond = 3 (auto): current choice
    int pc_auto_small = 0;
    // solver workspace
    double *krylov = nullptr;    // (restart+1) vectors
    int krylov_cap = 0;
    double *work[16] = {nullptr};
    int sm_count = 148;
    int max_smem = 232448;
};

// Host description of one input vector of the TMA-fed marcher (tma_march.cuh), built next
// to its VecRef (ksfd.cu: make_hvec): which buffers hold the owned and the ghost planes
struct TmaSrc {
    const double *base = nullptr;   // buffer of the owned planes
    long long base_fields = 0;      // fields (plane_pts doubles each) in that buffer
    int k0 = 0;                     // field index of plane 0
    const double *halo = nullptr;   // buffer of the ghost planes (nullptr: wrap, or in `base`)
    long long halo_fields = 0;
    int klo = 0, khi = 0;           // field index of plane -2 / plane nloc in the ghost buffer
    int wrap = 1;
    const unsigned long long *par = nullptr;    // device-side exchange counter (parity)
    int parshift = 0;               // fields between the two parity buffers
    // peer-memory exchange: flag words the neighbours publish, error words
    const volatile unsigned long long *flag_lo = nullptr, *flag_hi = nullptr;
    volatile int *err = nullptr;
    volatile unsigned long long *dead = nullptr;
};
struct HostVec {
    VecRef r;
    TmaSrc t;
};

// arguments of one Richardson sweep (sweep_op.cuh; ksfd.cu: sweep_solve_impl)
struct SweepHost {
    double *x, *rout;           // solution (updated in place), next residual
    double rsign;               // sign applied to the input residual (first sweep: sign of the rhs)
    int first;                  // x_0 = 0: x is written, not read
    int partial_cap;            // CTAs the partial-sum buffer has room for
    const void *fin;            // SweepFin (device-state pointers, options)
    // HaloPush or nullptr: the kernel also stores the boundary planes of rout into the
    // neighbours' ghost buffers and publishes the exchange (TMA-fed marcher only; the
    // caller then marks rout as pushed, else it launches the push kernel itself)
    const void *push;
};

// marching-kernel launchers (march_res.cu, march_jvp.cu, march_vel.cu); each is
// compiled once per dimension (-DKSFD_MARCH_DIM=2|3)
#define KSFD_DECL_MARCH(D)                                                            \
    int ksfd_march_residual_d##D(ksfd_ctx *c, const HostVec &u, const double *udot,   \
                                 const double *src, double *out, const void *push,    \
                                 cudaStream_t st);                                    \
    int ksfd_march_jvp_d##D(ksfd_ctx *c, const HostVec &coef, const HostVec &v,       \
                            const HostVec &pc, bool precond, double *out,             \
                            const int *skip, cudaStream_t st);                        \
    int ksfd_march_velocity_d##D(ksfd_ctx *c, const HostVec &u, double *vel,          \
                                 double *vmax, cudaStream_t st);                      \
    int ksfd_march_sweep_d##D(ksfd_ctx *c, const HostVec &coef, const HostVec &r,     \
                              const HostVec &pc, const SweepHost &a, const int *skip, \
                              cudaStream_t st);
KSFD_DECL_MARCH(2)
KSFD_DECL_MARCH(3)
void ksfd_free_plans(ksfd_ctx *c);
void ksfd_invalidate_plans(ksfd_ctx *c);
