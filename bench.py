#!/usr/bin/env python3
"""
bench.py — headline benchmark of the KSFD implicit time-stepping hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Workload (BASELINE.json configs[2]): 2-D 1024x1024 per-GPU tile, dof 3,
options84 physics (reference options84:20-46), h = 1/384, periodic, synthetic
random IC (rho = 9000 + 90 N(0,1), U = rho), fixed dt = 1e-3, ROSW ra34pw2,
one "step" = one implicit time step as the reference's loop does it
(KSFD/ksfdts.py:202-228): groom, TS.step (4 residuals + 1 Jacobian set-up +
4 linear solves), CFL velocity max.  For N > 1 the tile is fixed per GPU (weak
scaling): global grid 1024 x (1024 N), slab-decomposed along y with an NCCL
halo ring; value = grid-point-steps per second of the whole job.

Prints ONE JSON line (rank 0).  Extra keys: roofline (dominant kernel = the
Richardson sweep, or the fused J.v with GMRES), residual/jvp/sweep kernel numbers at
1024^2 and 256^3, cpu_baseline, e2e, clocks, gpu_launches.

CPU legs (cpu_baseline at N = 1, and --impl reference): the oracle's C restatement
(oracle/ksfd_oracle_c.c, OpenMP, every host thread, own process) on the SAME full
1024^2 configuration; falls back to the numpy port on 96^2 tiles if it cannot be built.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

TILE = 1024
H = 1.0 / 384
DT = 1e-3
KSP_RTOL = 1e-8
SEED = 793817931
ROSW_GAMMA = 0.435866521508459
ALG_BYTES_PER_PT = 72.0          # read u/v + read udot/coef-equivalent + write
ALG_BYTES_SWEEP = 120.0          # Richardson sweep: read u_lin, r, x; write x, r_new (5 x 24 B)
# stage-system solver: KSFD_BENCH_KSP=gmres|richardson|auto (default auto: the library's choice)
KSP_TYPE = os.environ.get('KSFD_BENCH_KSP', 'auto')
SOLVER_DESC = {'gmres': 'GMRES(30)',
               'richardson': 'stationary block-Jacobi sweeps fused into the stencil kernel',
               'auto': 'stationary block-Jacobi sweeps fused into the stencil kernel, '
                       'GMRES(30) fallback (not taken at this dt)'}


def phys_dict(dim, n):
    from helpers import phys84
    return phys84(dim, n, H)


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p))['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


# ---------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index = index
        self.window = None          # (t0, t1): only samples inside count
        self.samples = []
        self.stop = threading.Event()
        self.th = None

    def _run_nvml(self):
        """fast path: NVML in process (nvidia-ml-py), one sample every 10 ms"""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        R = [(nv.nvmlClocksThrottleReasonHwSlowdown, 0), (nv.nvmlClocksThrottleReasonHwThermalSlowdown, 1),
             (nv.nvmlClocksThrottleReasonSwThermalSlowdown, 2), (nv.nvmlClocksThrottleReasonSwPowerCap, 3)]
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self.stop.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            row = [str(sm), str(mx)] + ['Not Active'] * 4
            for bit, k in R:
                if bits & bit:
                    row[2 + k] = 'Active'
            self.samples.append(row + [time.perf_counter()])
            self.stop.wait(0.005)

    def _run(self):
        try:
            return self._run_nvml()
        except Exception:
            pass
        while not self.stop.is_set():
            try:
                o = subprocess.run(
                    ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                     '--format=csv,noheader,nounits'],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([x.strip() for x in o.split(',')] + [time.perf_counter()])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        rows = self.samples
        if self.window:
            inside = [r for r in rows if self.window[0] <= r[-1] <= self.window[1]]
            rows = inside or rows
        for s in rows:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
            except Exception:
                continue
            for nme, v in zip(names, s[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(nme)
        return dict(sm_mhz=float(np.median(sm)) if sm else None,
                    sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------
# CPU baseline: the oracle port (numpy + SuperLU) on a bounded sample
# ---------------------------------------------------------------------------
def _cpu_worker(args):
    n, nsteps, seed = args
    from helpers import oracle_physics, random_state
    from oracle import ksfd_oracle as O
    p = phys_dict(2, (n, n))
    ph = oracle_physics(p)
    u = random_state(p, seed, rel=0.0).reshape(ph.Vshape, order='F')
    t0 = time.perf_counter()
    t = 0.0
    for k in range(nsteps):
        u = O.groom(u, ph)
        u, _, _ = O.rosw_step(u, t, DT, ph)
        O.cfl_maxh(u, ph)
        t += DT
    return time.perf_counter() - t0


def cpu_baseline(nsteps=1, n=96, procs=1):
    """Oracle port timed on `procs` host cores (independent replicas, the way
    mpiexec ranks would each own a tile; no halo cost charged)."""
    import multiprocessing as mp
    if procs > 1:
        with mp.get_context('fork').Pool(procs) as pool:
            times = pool.map(_cpu_worker, [(n, nsteps, 100 + i) for i in range(procs)])
    else:
        times = [_cpu_worker((n, nsteps, 100))]
    wall = max(times)
    return dict(value=procs * n * n * nsteps / wall / 1e6, unit='Mpts*steps/s',
                cores=procs, kind='port',
                sample='%d ROSW step(s) of a %dx%d tile per core (oracle numpy '
                       '+ scipy SuperLU/MMD instead of MUMPS, same physics/h/dt), %d independent tiles'
                       % (nsteps, n, n, procs),
                seconds=wall)


def _host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def c_leg(nsteps, warm, n=TILE):
    """Runs in its own process (`bench.py --cpu-leg K,W`): the oracle's C restatement
    (oracle/ksfd_oracle_c.c, OpenMP on all host threads) on the SAME configuration as the
    GPU arm — the full n x n grid, options84 physics, h = 1/384, dt = 1e-3, one step = clamp,
    ROSW ra34pw2 step (4 residuals, Jacobian set-up, 4 stage solves to rtol 1e-8), CFL
    maxima.  Stage solves: point-block-Jacobi Richardson sweeps (the same iteration the GPU
    arm runs) instead of the reference's MUMPS LU, which at 3.1 M unknowns takes minutes per
    step — the substitution favours the CPU."""
    from helpers import oracle_physics, random_state
    from oracle import ksfd_oracle_c as OC
    p = phys_dict(2, (n, n))
    ph = oracle_physics(p)
    c = OC.COracle(ph)
    u = np.ascontiguousarray(random_state(p, 100, rel=0.0))
    its = 0
    for _ in range(warm):
        c.ts_step(u, DT, rtol=KSP_RTOL)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        _, k = c.ts_step(u, DT, rtol=KSP_RTOL)
        its += k
    wall = time.perf_counter() - t0
    # operator numbers on the same grid (best of 3)
    v = np.random.default_rng(8).standard_normal(u.size)
    c.jvp_setup(u, 1.0 / (ROSW_GAMMA * DT))
    ops = {}
    for name, fn in (('residual', lambda: c.dfdt(u)), ('jvp', lambda: c.jvp(v))):
        best = 1e30
        for _ in range(4):
            t1 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t1)
        ops[name] = dict(seconds=best, mpts_per_s=n * n / best / 1e6)
    finite = bool(np.isfinite(u).all())
    c.close()
    return dict(seconds=wall, steps=nsteps, warmup=warm, n=n, cores=OC.threads(),
                its_per_step=its / max(nsteps, 1), finite=finite, operators=ops)


def cpu_baseline_c(nsteps, warm=1, n=TILE, timeout=1500):
    """cpu_baseline from the C oracle in a fresh process with every host thread (torchrun
    sets OMP_NUM_THREADS=1 for its ranks; the leg gets its own environment)."""
    env = dict(os.environ)
    nth = int(os.environ.get('KSFD_CPU_THREADS', '0')) or _host_threads()
    env['OMP_NUM_THREADS'] = str(nth)
    env.setdefault('OMP_PROC_BIND', 'false')
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        env.pop(k, None)
    o = subprocess.run([sys.executable, os.path.abspath(__file__), '--cpu-leg',
                        '%d,%d,%d' % (nsteps, warm, n)],
                       capture_output=True, text=True, timeout=timeout, env=env)
    if o.returncode != 0:
        raise RuntimeError('cpu leg failed: ' + o.stderr[-500:])
    r = json.loads(o.stdout.strip().splitlines()[-1])
    if not r['finite']:
        raise RuntimeError('cpu leg produced non-finite values')
    val = n * n * r['steps'] / r['seconds'] / 1e6
    return dict(value=val, unit='Mpts*steps/s', cores=r['cores'], kind='port',
                sample='%d ROSW step(s) of the FULL %dx%d grid (same configuration as the GPU arm: '
                       'physics, h, dt, rtol) after %d warm-up step(s); oracle C restatement '
                       '(oracle/ksfd_oracle_c.c, OpenMP, %d threads); stage solves by point-block-Jacobi '
                       'Richardson sweeps (%.1f per step) instead of MUMPS LU'
                       % (r['steps'], n, n, r['warmup'], r['cores'], r['its_per_step']),
                seconds=r['seconds'], ms_per_step=1e3 * r['seconds'] / max(r['steps'], 1),
                ksp_its_per_step=r['its_per_step'], same_config=True,
                operators=r['operators'])


def cpu_operator_timing(n=TILE, reps=2):
    """Same-config CPU operator numbers: the oracle's residual f(u) and
    matrix-free J.v (numpy, one core) on the FULL 1024x1024 grid of the GPU roofline
    kernels.  (The reference's own Derivatives.dfdt runs through oracle/refharness only
    where /root/reference exists, which is not the case on the bench box: kind 'port'.)"""
    from helpers import oracle_physics, random_state
    from oracle import ksfd_oracle as O
    p = phys_dict(2, (n, n))
    ph = oracle_physics(p)
    u = random_state(p, 7, rel=0.0)
    v = np.random.default_rng(8).standard_normal(u.size)
    shift = 1.0 / (ROSW_GAMMA * DT)
    out = {}
    for name, fn in (('residual', lambda: O.dfdt(u, ph)),
                     ('jvp', lambda: O.jvp(u, v, shift, ph))):
        fn()
        best = 1e30
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        out[name] = dict(seconds=best, mpts_per_s=n * n / best / 1e6)
    return dict(grid='%dx%d' % (n, n), cores=1, kind='port',
                sample='oracle numpy residual f(u) and matrix-free J.v on the full grid, best of %d' % reps,
                **out)


def reference_arm(args):
    """CPU arm on the GPU arm's configuration: the oracle's C restatement on every host
    thread, K steps of the full 1024^2 grid (at N > 1 the sample stays ONE per-GPU tile of the
    weak-scaling grid: CPU throughput in points*steps/s does not depend on the number of
    tiles, and K steps must end within minutes).  The reference itself (PETSc + MUMPS over
    MPI) cannot be installed in this image: kind 'port'."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    t0 = time.perf_counter()
    direct = None
    try:
        cb = cpu_baseline_c(args.steps, max(args.warmup, 0))
        same = max(args.gpus, 1) == 1
        par = '%d host threads (OpenMP), one process' % cb['cores']
        how = ('CPU arm: oracle C restatement, stage solves by point-block-Jacobi Richardson sweeps '
               '(%.1f per step) instead of MUMPS LU' % cb['ksp_its_per_step'])
        if not same:
            how += ('; each step a bounded sample of the %d-tile grid = ONE %dx%d tile (CPU throughput in '
                    'points*steps/s does not depend on the number of tiles)' % (args.gpus, TILE, TILE))
        try:        # what a direct solver costs: one step of a 96^2 tile, numpy port + SuperLU
            d = cpu_baseline(1, 96, 1)
            direct = dict(value=d['value'], unit=d['unit'], cores=1, sample=d['sample'],
                          seconds=d['seconds'])
        except Exception as e:          # noqa: BLE001
            direct = dict(error=str(e)[:200])
    except Exception as e:              # noqa: BLE001 — the numpy port is always there
        sys.stderr.write('bench.py: C oracle leg unavailable (%s); numpy port on 96^2 tiles\n' % e)
        procs = min(os.cpu_count() or 1, 32)
        n = 96 if args.steps <= 20 else 64
        cb = cpu_baseline(args.steps, n, procs)
        cb['ms_per_step'] = 1e3 * cb['seconds'] / max(args.steps, 1)
        same = False
        par = '%d host cores, independent tiles' % procs
        how = ('CPU arm: numpy port, direct LU (SuperLU), each step a bounded sample = one '
               '%dx%d tile per core' % (n, n))
    wall = time.perf_counter() - t0
    line = dict(impl='reference', metric='implicit TS throughput (ROSW steps x grid points)',
                value=cb['value'], unit='Mpts*steps/s', n_gpus=args.gpus,
                steps=args.steps, warmup=args.warmup,
                ms_per_step=cb['ms_per_step'],
                higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f64', data='synthetic',
                config=dict(workload='2-D %dx%d tile per GPU (global %dx%d), dof 3, options84 physics, '
                                     'h=1/384, dt=1e-3, ROSW ra34pw2, rtol %.0e; %s'
                                     % (TILE, TILE, TILE, TILE * max(args.gpus, 1), KSP_RTOL, how),
                            parallelism=par),
                cpu_baseline=dict(value=cb['value'], unit=cb['unit'], cores=cb['cores'],
                                  kind='port', sample=cb['sample']),
                e2e=dict(value=cb['value'], unit='Mpts*steps/s',
                         h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                same_config=same,
                # operator numbers on the same full grid: C restatement on all threads, and the
                # numpy port on one core
                cpu_operator_1024x1024=dict(cpu_operator_timing(),
                                            c_all_threads=cb.get('operators')),
                cpu_direct_solver_sample=direct,
                wall_s=wall)
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------
def time_kernel(fn, nrot, reps, warm=3):
    """Average CUDA-event time (us) of fn(i) over reps launches, rotating over
    nrot buffer sets so that successive launches read cold (non-L2) data."""
    import torch
    for i in range(warm):
        fn(i % nrot)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i % nrot)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def kernel_rooflines(dim, n, reps=20):
    """residual and J.v kernel throughput on one GPU for grid n."""
    import torch
    from helpers import product_physics
    from ksfd_b200 import core
    peak, src = measured_peak()
    p = phys_dict(dim, n)
    ctx = core.Context(dim, n, 3)
    ctx.set_physics(product_physics(p))
    npts, N = ctx.npts, ctx.npts * 3
    nrot = max(2, int(2.5 * 126e6 // (N * 8 * 3)) + 1)
    gen = torch.Generator(device='cuda').manual_seed(SEED)
    us = [(9000 + 90 * torch.randn(npts, generator=gen, device='cuda',
                                   dtype=torch.float64)).repeat_interleave(3).contiguous()
          for _ in range(nrot)]
    vs = [torch.randn(N, generator=gen, device='cuda', dtype=torch.float64)
          for _ in range(nrot)]
    outs = [torch.empty(N, device='cuda', dtype=torch.float64) for _ in range(nrot)]
    ctx.jvp_setup(us[0], 1.0 / (ROSW_GAMMA * DT))
    out = {}
    # the sweep updates x in place: the u buffers (not read by the J.v-side kernels once the
    # linearisation is set up) serve as its rotating x operands
    for name, fn, bpp in (('residual', lambda i: ctx.residual(us[i], vs[i], None, outs[i]), ALG_BYTES_PER_PT),
                          ('jvp', lambda i: ctx.jvp(vs[i], outs[i]), ALG_BYTES_PER_PT),
                          ('jvp_precond', lambda i: ctx.jvp(vs[i], outs[i], precond=True), ALG_BYTES_PER_PT),
                          ('sweep', lambda i: ctx.sweep(vs[i], us[i], outs[i], norms=False), ALG_BYTES_SWEEP)):
        if name == 'sweep':
            ctx.residual(us[0], vs[0], None, outs[0])       # (the residual above needed us intact)
        us_ = time_kernel(fn, nrot, reps)
        ach = npts * bpp / (us_ * 1e-6) / 1e9
        out[name] = dict(us=us_, gpts_per_s=npts / us_ / 1e3, achieved_gbs=ach,
                         frac=ach / peak, algorithmic_bytes_per_point=bpp)
    ctx.close()
    del us, vs, outs
    torch.cuda.empty_cache()
    return out, peak, src


def step_timing_3d(n=256, nsteps=5, warm=2):
    """implicit ROSW steps of the 3-D n^3 problem on one GPU (BASELINE configs[3] tile):
    the same step as the headline, reported next to it."""
    import torch
    from helpers import product_physics
    from ksfd_b200 import core
    p = phys_dict(3, (n, n, n))
    ctx = core.Context(3, (n, n, n), 3)
    ctx.set_physics(product_physics(p))
    rng = np.random.default_rng(np.random.SeedSequence(SEED).spawn(1)[0])
    rho = 9000.0 + 90.0 * rng.standard_normal(ctx.npts)
    u = ctx.to_internal(torch.from_numpy(np.repeat(rho, 3)).cuda())
    opts = core.ts_options(ts_type='rosw', adapt='none', atol=0.01, rtol=1e-6,
                           ksp_rtol=KSP_RTOL, ksp_max_it=2000, restart=30, ksp_type=KSP_TYPE,
                           groom=True, velocity_max=True)
    t, its = 0.0, 0
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    for k in range(warm + nsteps):
        if k == warm:
            torch.cuda.synchronize()
            e0.record()
            its = 0
        r = ctx.ts_step(u, t, DT, opts)         # clamp + step + CFL maxima (flags in opts)
        if not r.accepted or not r.have_vmax:
            raise RuntimeError('3-D time step failed')
        t = r.t_new
        its += r.ksp_its
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / nsteps
    ctx.close()
    torch.cuda.empty_cache()
    return dict(ms_per_step=ms, steps_per_sec=1e3 / ms, mpts_steps_per_s=n ** 3 / ms / 1e3,
                gmres_its_per_step=its / nsteps, steps=nsteps)


def build_ctx(dim, n, local, rank, world):
    from helpers import product_physics
    from ksfd_b200 import core
    ctx = core.Context(dim, n, 3, device=local, rank=rank, nranks=world)
    ctx.set_physics(product_physics(phys_dict(dim, n)))
    if world > 1:
        from ksfd_b200 import parallel
        parallel.init_comm(ctx)
    return ctx


def synthetic_state(ctx, rank, world):
    """rho = 9000 + 90 N(0,1) from the rank's spawned stream, U = rho (internal layout)"""
    import torch
    rng = np.random.default_rng(np.random.SeedSequence(SEED).spawn(world)[rank])
    rho = 9000.0 + 90.0 * rng.standard_normal(ctx.npts)
    host = torch.from_numpy(np.repeat(rho, 3)).pin_memory()
    return host, ctx.to_internal(host.cuda())


def timed_steps(ctx, u, opts, nsteps, warm, world):
    """ms per step (max over ranks), GMRES iterations per step"""
    import torch
    import torch.distributed as dist
    t, its = 0.0, 0
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    for k in range(warm + nsteps):
        if k == warm:
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0.record()
            its = 0
        r = ctx.ts_step(u, t, DT, opts)         # clamp + step + CFL maxima (flags in opts)
        if not r.accepted or not r.have_vmax:
            raise RuntimeError('time step failed (ksp_fail=%d)' % r.ksp_fail)
        t = r.t_new
        its += r.ksp_its
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()) / nsteps, its / nsteps


def strong_scaling_records(local, rank, world, opts, nsteps=5, warm=2):
    """BASELINE configs[2] / configs[3] inside the one bench line: the fixed GLOBAL grids
    1024^2 and 256^3 split into `world` slabs (all ranks), and — on rank 0 alone — the same
    grids on one GPU, so that the strong-scaling efficiency is measured in this run."""
    import torch
    import torch.distributed as dist
    out = {}
    for key, dim, n in (('strong_1024x1024', 2, (TILE, TILE)),
                        ('strong_256x256x256', 3, (256, 256, 256))):
        ctx = build_ctx(dim, n, local, rank, world)
        _, u = synthetic_state(ctx, rank, world)
        ms, its = timed_steps(ctx, u, opts, nsteps, warm, world)
        rows = ctx.last_count
        ctx.close()
        del u
        torch.cuda.empty_cache()
        rec = dict(n_gpus=world, ms_per_step=ms, steps_per_sec=1e3 / ms,
                   gmres_its_per_step=its, planes_per_gpu=rows, steps=nsteps)
        if world > 1:
            if rank == 0:
                c1 = build_ctx(dim, n, local, 0, 1)
                _, u1 = synthetic_state(c1, 0, 1)
                ms1, its1 = timed_steps(c1, u1, opts, nsteps, warm, 1)
                c1.close()
                del u1
                torch.cuda.empty_cache()
                rec.update(ms_per_step_1gpu=ms1, speedup=ms1 / ms,
                           efficiency=ms1 / ms / world, gmres_its_per_step_1gpu=its1)
            dist.barrier()
        out[key] = rec
    return out


def multi_gpu_parity(ctx, u, local, rank, world, n, opts):
    """Correctness of the distributed run inside the bench line (VERDICT r1 #3b): residual,
    J.v, fused A*M^-1 v and one ROSW step of the slab-decomposed problem against a
    single-GPU recomputation of the same global problem on rank 0."""
    import torch
    import torch.distributed as dist
    gen = torch.Generator(device='cuda').manual_seed(SEED + 17 + rank)
    N = ctx.npts * 3
    ud = torch.randn(N, generator=gen, device='cuda', dtype=torch.float64)
    v = torch.randn(N, generator=gen, device='cuda', dtype=torch.float64)
    shift = 1.0 / (ROSW_GAMMA * DT)
    F = ctx.residual(u, ud)
    ctx.jvp_setup(u, shift)
    Jv = ctx.jvp(v)
    Jp = ctx.jvp(v, precond=True)
    u1 = u.clone()
    ctx.groom(u1)
    r = ctx.ts_step(u1, 0.0, DT, opts)
    mine = [u, ud, v, F, Jv, Jp, u1]

    def gather(t):
        parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t, parts, dst=0)
        return torch.cat(parts) if rank == 0 else None      # slabs of the last axis: planes are outermost

    full = [gather(t) for t in mine]
    rec = None
    if rank == 0:
        c1 = build_ctx(2, n, local, 0, 1)
        U, UD, V, Fm, Jvm, Jpm, U1m = full
        F1 = c1.residual(U, UD)
        c1.jvp_setup(U, shift)
        Jv1 = c1.jvp(V)
        Jp1 = c1.jvp(V, precond=True)
        U1 = U.clone()
        c1.groom(U1)
        r1 = c1.ts_step(U1, 0.0, DT, opts)
        rel = float(((U1m - U1).abs().max() / U1.abs().max()).item())
        rec = dict(residual_bit_identical=bool(torch.equal(Fm, F1)),
                   jvp_bit_identical=bool(torch.equal(Jvm, Jv1)),
                   jvp_precond_bit_identical=bool(torch.equal(Jpm, Jp1)),
                   rosw_step_max_rel_err=rel, rosw_step_tol=1e-12,
                   gmres_its=[int(r.ksp_its), int(r1.ksp_its)],
                   ok=bool(torch.equal(Fm, F1) and torch.equal(Jvm, Jv1)
                           and torch.equal(Jpm, Jp1) and rel < 1e-12),
                   reference='rank 0 recomputes the same global %dx%d problem on one GPU' % n)
        c1.close()
    dist.barrier()
    return rec


def native_arm(args):
    # stdout carries the ONE JSON line: NCCL's own log lines ("NCCL version ...",
    # NCCL_DEBUG=INFO output) go to stderr instead of their default, stdout
    os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
    import torch
    import torch.distributed as dist
    from helpers import product_physics
    from ksfd_b200 import _lib, core

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the native arm has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    # default: weak scaling, one 1024^2 tile per GPU; --workload selects the
    # strong-scaling cases of BASELINE configs[2] / configs[3] (fixed global grid)
    if args.workload == 'strong-1024':
        dim, n, scaling = 2, (TILE, TILE), 'strong'
        wl = '2-D 1024x1024 GLOBAL grid split into %d slab(s)' % world
    elif args.workload == 'strong-256':
        dim, n, scaling = 3, (256, 256, 256), 'strong'
        wl = '3-D 256^3 GLOBAL grid split into %d slab(s) along z' % world
    else:
        dim, n, scaling = 2, (TILE, TILE * world), 'weak'
        wl = '2-D 1024x1024 tile per GPU (global 1024x%d)' % n[1]
    p = phys_dict(dim, n)
    ctx = core.Context(dim, n, 3, device=local, rank=rank, nranks=world)
    ctx.set_physics(product_physics(p))
    if world > 1:
        from ksfd_b200 import parallel
        parallel.init_comm(ctx)
    if os.environ.get('KSFD_BENCH_UNDERPREDICT'):       # test knob: exercise the late-sweep path
        ctx.set_option('sweep_underpredict', int(os.environ['KSFD_BENCH_UNDERPREDICT']))
    # synthetic IC, per-rank stream as the reference spawns it
    # (KSFD/ksfdrandom.py:46-49)
    rng = np.random.default_rng(np.random.SeedSequence(SEED).spawn(world)[rank])
    rho = 9000.0 + 90.0 * rng.standard_normal(ctx.npts)
    u_host = torch.from_numpy(np.repeat(rho, 3)).pin_memory()   # reference layout
    u_ref = u_host.cuda()
    u = ctx.to_internal(u_ref)                                  # internal layout
    opts = core.ts_options(ts_type='rosw', adapt='none', atol=0.01, rtol=1e-6,
                           ksp_rtol=KSP_RTOL, ksp_max_it=2000, restart=30, ksp_type=KSP_TYPE,
                           groom=True, velocity_max=True)
    state = dict(t=0.0, its=0)

    def step():
        # clamp, ROSW step and CFL maxima — the body of the reference's step loop
        # (KSFD/ksfdts.py:205-227) — in ONE library call, as ksfd_b200.ts.KSFDTS.solve makes it
        r = ctx.ts_step(u, state['t'], DT, opts)
        if not r.accepted or not r.have_vmax:
            raise RuntimeError('time step failed (ksp_fail=%d)' % r.ksp_fail)
        state['t'] = r.t_new
        state['its'] += r.ksp_its
        return [r.vmax[i] for i in range(dim)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the clock sampler (NVML, every 5 ms) runs from before the warm-up; only the
    # samples taken inside the timed region are reported
    with ClockSampler(local) as cs:
        for _ in range(args.warmup):
            step()
        # ---- device-resident timing -------------------------------------
        barrier()
        state['its'] = 0
        l0 = _lib.launch_count()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        cs.window = (w0, time.perf_counter())
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = _lib.launch_count() - l0
    its_per_step = state['its'] / max(args.steps, 1)
    gpts = int(np.prod(n))
    value = gpts * args.steps / (ms * 1e-3) / 1e6
    # ---- end-to-end through host buffers --------------------------------
    # two pinned host buffers, ping-pong: each step reads its input from one
    # (H2D) and returns its result into the other (D2H); no extra host copy
    hbuf = [u_host, torch.empty_like(u_host).pin_memory()]
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    esteps = max(1, min(args.steps, 10))
    for i in range(esteps):
        src, dst = hbuf[i % 2], hbuf[(i + 1) % 2]
        u_ref.copy_(src, non_blocking=True)         # H2D of the step's input
        ctx.to_internal(u_ref, out=u)               # reference -> internal layout
        step()
        ctx.from_internal(u, out=u_ref)
        dst.copy_(u_ref, non_blocking=True)         # D2H of the step's result
        torch.cuda.synchronize()                    # the host owns the result
    t1.record()
    barrier()
    ems = torch.tensor([t0.elapsed_time(t1)], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    e2e_val = gpts * esteps / (float(ems.item()) * 1e-3) / 1e6
    nbytes = ctx.npts * 3 * 8
    ctx_npts_local = ctx.npts
    # ---- the stencil kernels INSIDE the step loop (CUDA events around every launch; a
    # separate pass, not the timed one: the events cost a little on the stream)
    ctx.set_option('profile', 1)
    for _ in range(3):
        step()
    prof = ctx.profile_fetch()
    ctx.set_option('profile', 0)
    # ---- several GPUs: correctness of the distributed run in this very line
    parity = None
    if world > 1 and args.workload == 'weak-1024':
        parity = multi_gpu_parity(ctx, u, local, rank, world, n, opts)
    ctx.close()
    strong = None
    if not args.quick and args.workload == 'weak-1024' and world > 1:
        del u
        torch.cuda.empty_cache()
        strong = strong_scaling_records(local, rank, world, opts)

    extra = {}
    roof = None
    cpu = None
    if rank == 0:
        torch.cuda.empty_cache()
        k2, peak, psrc = kernel_rooflines(2, (TILE, TILE))
        extra['kernels_1024x1024'] = k2
        if not args.quick:
            k3, _, _ = kernel_rooflines(3, (256, 256, 256), reps=10)
            extra['kernels_256x256x256'] = k3
            if world == 1:
                extra['step_256x256x256'] = step_timing_3d()
            if world == 1:
                # BASELINE configs[4]: per-GPU tiles 512^2 .. 4096^2
                sweep = {'1024x1024': k2, '256x256x256': k3}
                for m in (512, 2048, 4096):
                    sweep['%dx%d' % (m, m)], _, _ = kernel_rooflines(2, (m, m))
                extra['kernel_sweep'] = sweep
        # the dominant kernel of the step: the Richardson sweep (fused A*M^-1 stencil +
        # x / r update + norms) when the stage solves ran on sweeps, else the fused J.v
        swept = prof['sweep_launches'] > prof['jvp_launches']
        domkey, dombytes = ('sweep', ALG_BYTES_SWEEP) if swept else ('jvp_precond', ALG_BYTES_PER_PT)
        dom = k2[domkey]
        traffic, tsrc = None, None
        tp = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get(domkey + '_1024x1024_bytes_per_launch')
            tsrc = tj.get(domkey + '_source', tj.get('source'))
        tmatch = None
        if os.path.exists(tp) and tj.get('kernel_sources'):
            import hashlib
            hh = hashlib.sha256()
            for f in tj['kernel_sources']:
                hh.update(open(os.path.join(ROOT, 'ksfd_b200', 'csrc', f), 'rb').read())
            tmatch = hh.hexdigest() == tj.get('kernel_sources_sha256_at_capture')
        roof = dict(bound='hbm',
                    kernel='k_tma_march<2,256,1,SweepOp<2,2>> (Richardson sweep: fused A*M^-1 stencil, '
                           'x and r update, norms; TMA-fed)' if swept else
                           'k_tma_march<2,256,1,JvpOp<2,2,precond>> (fused A*M^-1 v, TMA-fed)',
                    achieved=dom['achieved_gbs'], peak=peak, unit='GB/s',
                    frac=dom['frac'], traffic=traffic, traffic_source=tsrc, traffic_matches_build=tmatch,
                    peak_source=psrc,
                    algorithmic_bytes_per_point=dombytes,
                    algorithmic_bytes='sweep: read u_lin, r, x; write x, r_new = 5 x 24 B/point'
                                      if swept else 'J.v: read u_lin, v; write out = 3 x 24 B/point',
                    points_per_launch=TILE * TILE, us_per_launch=dom['us'],
                    timing='CUDA events, 20 launches rotating over buffer sets > 2.5x L2 (cold)')
        pk = 'sweep' if swept else 'jvp'
        if prof[pk + '_launches'] > 0:
            # the same kernel where it actually runs: inside the ROSW step loop, its
            # operands partly L2-resident (coefficient field 42 MB, vectors 25 MB)
            us_in = 1e3 * prof[pk + '_ms'] / prof[pk + '_launches']
            ach = ctx_npts_local * dombytes / (us_in * 1e-6) / 1e9
            roof['in_step'] = dict(us_per_launch=us_in, achieved=ach, frac=ach / peak,
                                   launches=prof[pk + '_launches'],
                                   launches_incl_skipped=prof[pk + '_launches_all'],
                                   points_per_launch=ctx_npts_local,
                                   timing='CUDA events around every launch of the kernel in 3 ROSW '
                                          'steps (separate pass after the timed region)')
        if prof['residual_launches'] > 0:
            us_r = 1e3 * prof['residual_ms'] / prof['residual_launches']
            ach = ctx_npts_local * ALG_BYTES_PER_PT / (us_r * 1e-6) / 1e9
            extra['residual_in_step'] = dict(us_per_launch=us_r, achieved_gbs=ach, frac=ach / peak,
                                             launches=prof['residual_launches'])
        # where the step goes: mean device time of the active launches of each kind
        extra['in_step_kernel_us'] = {
            nm: dict(us=1e3 * prof[nm + '_ms'] / max(prof[nm + '_launches'], 1),
                     launches_per_step=prof[nm + '_launches'] / 3.0,
                     ms_per_step=prof[nm + '_ms_all'] / 3.0)
            for nm in ('sweep', 'jvp', 'residual', 'mdot', 'orth', 'first_vector', 'cycle_begin')}
        if world == 1 and not args.no_cpu:
            try:        # same configuration, every host thread (C restatement of the oracle)
                cpu = cpu_baseline_c(20, 1)
                ops = cpu.pop('operators')
            except Exception as e:      # noqa: BLE001 — the numpy port is always there
                sys.stderr.write('bench.py: C oracle leg unavailable (%s); numpy port\n' % e)
                cpu, ops = cpu_baseline(2, 96, 1), None
            extra['cpu_operator_1024x1024'] = dict(cpu_operator_timing(), c_all_threads=ops)
    if strong is not None:
        extra.update(strong)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    line = dict(metric='implicit TS throughput (ROSW steps x grid points)',
                value=value, unit='Mpts*steps/s', n_gpus=world, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms / args.steps,
                higher_is_better=True, scaling=scaling, vs_baseline=None,
                dtype='f64', data='synthetic',
                steps_per_sec=args.steps / (ms * 1e-3),
                config=dict(workload='%s, dof 3, '
                                     'options84 physics, h=1/384, dt=1e-3, ROSW ra34pw2, '
                                     'stage solves: %s, point-block Jacobi, rtol %.0e'
                                     % (wl, SOLVER_DESC[KSP_TYPE], KSP_RTOL),
                            ksp_type=KSP_TYPE,
                            parallelism='slab%d' % world,
                            l2='step working set (coefficient field 42 MB, 2 residual + 10 stage '
                               'vectors of 25 MB; GMRES: + 31 Krylov vectors) exceeds the 126 MB L2; kernel-only timings rotate '
                               'over >2.5x L2 of distinct buffers'),
                gmres_its_per_step=its_per_step, ksp_its_per_step=its_per_step,
                gpu_launches=int(launches),
                clocks=cs.summary(),
                e2e=dict(value=e2e_val, unit='Mpts*steps/s',
                         h2d_bytes_per_step=nbytes, d2h_bytes_per_step=nbytes,
                         steps=esteps),
                roofline=roof, cpu_baseline=cpu, parity=parity, **extra)
    print(json.dumps(line), flush=True)


def main():
    wd = os.environ.get('KSFD_BENCH_WATCHDOG')
    if wd:      # debugging aid: dump all Python stacks and exit after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(wd), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--workload', default='weak-1024',
                    choices=['weak-1024', 'strong-1024', 'strong-256'],
                    help='weak-1024 (default, the headline): one 1024^2 tile per GPU; '
                         'strong-1024 / strong-256: BASELINE configs[2] / configs[3], the '
                         'global 1024^2 / 256^3 grid split over the GPUs')
    ap.add_argument('--quick', action='store_true', help='skip the 256^3 kernel timings')
    ap.add_argument('--no-cpu', action='store_true', help='skip the CPU baseline')
    ap.add_argument('--cpu-leg', default=None, help=argparse.SUPPRESS)     # internal: K,W,n
    ap.add_argument('--sweep', action='store_true',
                    help='kernel rooflines over per-GPU tiles 512^2..4096^2 and 256^3 '
                         '(BASELINE configs[4]); prints one JSON object, not the bench line')
    args = ap.parse_args()
    if args.cpu_leg:
        k, w, n = (int(x) for x in args.cpu_leg.split(','))
        print(json.dumps(c_leg(k, w, n)), flush=True)
        return
    if args.sweep:
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        out = {}
        for n in (512, 1024, 2048, 4096):
            out['%dx%d' % (n, n)], peak, _ = kernel_rooflines(2, (n, n))
        out['256x256x256'], peak, _ = kernel_rooflines(3, (256, 256, 256), reps=10)
        print(json.dumps(dict(sweep=out, peak_gbs=peak, algorithmic_bytes_per_point=ALG_BYTES_PER_PT)))
        return
    if args.impl == 'reference':
        reference_arm(args)
    else:
        native_arm(args)


if __name__ == '__main__':
    main()
