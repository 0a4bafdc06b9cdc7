"""
Pin the oracle's C restatement (oracle/ksfd_oracle_c.c, OpenMP) against
  (a) the golden vectors produced by the reference's own code (tests/golden, made by
      oracle/make_golden.py from the unmodified KSFD.Derivatives), and
  (b) the numpy oracle (oracle/ksfd_oracle.py), operation by operation,
and its iterative stage solves against SuperLU.  CPU only.  The C oracle is the CPU baseline
bench.py times on the full 1024^2 grid (cpu_baseline / --impl reference).
"""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from helpers import (check_field, cond_scale, golden_names, load_golden, oracle_physics, phys84,
                     random_state, relerr)
from oracle import ksfd_oracle as O
from oracle import ksfd_oracle_c as OC

NAMES = golden_names()
TOL_F = 5e-12        # as tests/test_oracle_vs_golden.py
TOL_J = 1e-11


def test_library_abi():
    L = OC.lib()
    assert L.oc_abi() == 1
    assert OC.threads() >= 1
    for sym in ('oc_create', 'oc_destroy', 'oc_groom', 'oc_dfdt', 'oc_ifunction', 'oc_velocity',
                'oc_velocity_max', 'oc_jvp_setup', 'oc_jvp', 'oc_pc_apply', 'oc_get_minv',
                'oc_solve', 'oc_rosw_step', 'oc_beuler_step', 'oc_get_stage', 'oc_ts_step'):
        assert hasattr(L, sym), sym


@pytest.mark.parametrize('name', NAMES)
def test_c_oracle_vs_reference_goldens(name):
    """residual (with sources), velocity and J.v against what the reference's own code produced"""
    g, physs, nrec = load_golden(name)
    for r in range(nrec):
        ph = oracle_physics(physs[r])
        c = OC.COracle(ph)
        u = g['u_%d' % r]
        f = c.dfdt(u, sources=g['src_%d' % r])
        bad = check_field(f, g['f_%d' % r], ph.dof, TOL_F, cond_scale(ph, u))
        assert bad < 1.0, (name, r, bad)
        v = c.velocity(u)
        vr = g['vel_%d' % r]
        assert relerr(v, vr) < TOL_F or np.abs(v - vr).max() < 1e-13, (name, r)
        c.jvp_setup(u, 0.0)
        for m in range(3):
            mf = -c.jvp(g['v_%d' % r][m])
            assert relerr(mf, g['Jv_%d' % r][m], ph.dof) < TOL_J, (name, r, m)
        c.close()


CASES = [(1, (40,)), (2, (24, 20)), (2, (7, 33)), (3, (8, 6, 10))]


@pytest.mark.parametrize('dim,n', CASES)
def test_c_oracle_vs_numpy_oracle(dim, n):
    p = phys84(dim, n)
    ph = oracle_physics(p)
    c = OC.COracle(ph)
    u = random_state(p, 3)
    rng = np.random.default_rng(4)
    ud = rng.standard_normal(u.size)
    v = rng.standard_normal(u.size)
    cond = cond_scale(ph, u)
    F = c.ifunction(u, ud)
    Fr = O.ifunction(u, ud, ph).reshape(-1, order='F')
    assert check_field(F, Fr, ph.dof, 1e-13, cond, ncond=8.0) < 1.0
    shift = 1.0 / (O.ROSW_GAMMA * 1e-3)
    c.jvp_setup(u, shift)
    assert relerr(c.jvp(v), O.jvp(u, v, shift, ph).reshape(-1, order='F'), ph.dof) < 1e-14
    # M^-1 = inverse of the oracle's assembled diagonal blocks
    B = O.block_diagonal(u, shift, ph)
    Mi = c.minv()
    eye = np.eye(ph.dof)
    assert max(np.abs(Mi[i] @ B[i] - eye).max() for i in range(ph.npts)) < 1e-12
    z = c.pc_apply(v)
    zr = np.einsum('pab,pb->pa', Mi, v.reshape(ph.npts, ph.dof)).ravel()
    assert relerr(z, zr) < 1e-14
    vm = c.velocity_max(u)
    vr = np.abs(O.velocity(u, ph)).reshape(dim, -1).max(axis=1)
    assert np.allclose(vm, vr, rtol=1e-11, atol=0)
    c.close()


def test_groom_nan_and_negative():
    p = phys84(2, (6, 5))
    ph = oracle_physics(p)
    c = OC.COracle(ph)
    u = random_state(p, 1)
    u[0] = np.nan
    u[4] = -3.0
    u[7] = np.nan
    u[9] = 0.0
    ur = O.groom(u.copy().reshape(ph.Vshape, order='F'), ph).reshape(-1, order='F')
    assert np.array_equal(c.groom(u), ur)           # bit-exact
    # and the residual clamps its own copy the same way
    f = c.dfdt(u)
    fr = O.dfdt(u, ph).reshape(-1, order='F')
    assert check_field(f, fr, ph.dof, 1e-12, cond_scale(ph, ur)) < 1.0
    c.close()


@pytest.mark.parametrize('ksp', ['auto', 'richardson', 'gmres'])
def test_solve_vs_superlu(ksp):
    p = phys84(2, (20, 16))
    ph = oracle_physics(p)
    c = OC.COracle(ph)
    u = random_state(p, 5)
    b = np.random.default_rng(6).standard_normal(u.size)
    shift = 1.0 / (O.ROSW_GAMMA * 1e-3)
    c.jvp_setup(u, shift)
    x, info = c.solve(b, rtol=1e-12, ksp_type=ksp)
    xr = spla.splu(O.ijacobian(u, shift, ph).tocsc()).solve(b)
    assert info['converged'] and info['gmres'] == (ksp == 'gmres')
    assert relerr(x, xr) < 1e-10
    # the norm it reports is the true residual norm
    rn = np.linalg.norm(b - c.jvp(x))
    assert rn <= 1.05 * max(info['rnorm'], 1e-12 * info['bnorm'])
    c.close()


def test_large_step_falls_back_to_gmres():
    p = phys84(2, (20, 16))
    ph = oracle_physics(p)
    c = OC.COracle(ph)
    u = random_state(p, 5)
    b = np.random.default_rng(6).standard_normal(u.size)
    shift = 1.0 / (O.ROSW_GAMMA * 1.0)
    c.jvp_setup(u, shift)
    x, info = c.solve(b, rtol=1e-10)
    assert info['converged'] and info['gmres']
    xr = spla.splu(O.ijacobian(u, shift, ph).tocsc()).solve(b)
    assert relerr(x, xr) < 1e-7
    c.close()


@pytest.mark.parametrize('dim,n,dt', [(1, (40,), 1e-3), (2, (16, 12), 1e-3), (2, (16, 12), 0.1),
                                      (3, (6, 5, 7), 1e-3)])
def test_rosw_step_vs_numpy_oracle(dim, n, dt):
    """iterative stage solves (rtol 1e-13) against the numpy oracle's SuperLU step"""
    p = phys84(dim, n)
    ph = oracle_physics(p)
    c = OC.COracle(ph)
    u = random_state(p, 2)
    un, ue, info = c.rosw_step(u, dt, rtol=1e-13)
    ur, uer, Y = O.rosw_step(u, 0.0, dt, ph)
    assert relerr(un, ur.reshape(-1, order='F')) < 1e-10
    assert relerr(ue, uer.reshape(-1, order='F')) < 1e-10
    for j in range(4):
        assert relerr(c.stage(j), Y[j].reshape(-1, order='F')) < 1e-8
    assert info['its'] > 0
    c.close()


def test_ts_step_is_groom_step_cfl():
    p = phys84(2, (16, 12))
    ph = oracle_physics(p)
    c = OC.COracle(ph)
    u = random_state(p, 9)
    u[0] = -1.0
    w = u.copy()
    vm, its = c.ts_step(w, 1e-3, rtol=1e-13)
    ug = O.groom(u.copy().reshape(ph.Vshape, order='F'), ph)
    ur, _, _ = O.rosw_step(ug, 0.0, 1e-3, ph)
    assert relerr(w, ur.reshape(-1, order='F')) < 1e-10
    vr = np.abs(O.velocity(ur, ph)).reshape(2, -1).max(axis=1)
    assert np.allclose(vm, vr, rtol=1e-8)
    c.close()


def test_bench_cpu_leg_runs_in_its_own_process():
    """bench.py's CPU leg (`--cpu-leg K,W,n`, what cpu_baseline / --impl reference time) on a
    small grid: one JSON line, finite state, every host thread even under torchrun's
    OMP_NUM_THREADS=1."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS='2')
    o = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--cpu-leg', '2,1,48'],
                       capture_output=True, text=True, timeout=300, env=env)
    assert o.returncode == 0, o.stderr[-400:]
    r = json.loads(o.stdout.strip().splitlines()[-1])
    assert r['finite'] and r['steps'] == 2 and r['n'] == 48 and r['cores'] == 2
    assert 20 <= r['its_per_step'] <= 40            # 4 solves of 7-8 sweeps at dt = 1e-3
    assert r['operators']['residual']['mpts_per_s'] > 0


def test_adaptive_trajectory_matches_numpy_oracle():
    """TSAdapt basic from a tiny first step (as the adaptive option files start): same accepted
    times, same states, while dt grows by more than three decades."""
    p = phys84(2, (16, 12))
    ph = oracle_physics(p)
    c = OC.COracle(ph)
    u0 = random_state(p, 12)
    adapt = dict(atol=1e-2, rtol=1e-6, clip=(0.1, 5.0))
    ref = O.integrate(u0, 0.0, 1e-6, 12, ph, adapt=adapt)
    got, rejected = OC.integrate(c, u0, 0.0, 1e-6, 12, adapt=adapt, rtol=1e-13)
    assert len(got) == len(ref) == 12
    for (tr, ur), (tg, ug) in zip(ref, got):
        assert abs(tr - tg) <= 1e-9 * tr
        assert relerr(ug, ur.reshape(-1, order='F')) < 1e-9
    assert ref[-1][0] > 1e-3            # dt really grew (from 1e-6)
    c.close()


def test_c_oracle_vs_numpy_oracle_random_physics():
    """random ligand groups (1-3 groups, up to 5 ligands), both cap potentials, odd extents down
    to the stencil width, all dimensions: residual, J.v and block inverses of the two oracles"""
    rng = np.random.default_rng(2024)
    for case in range(24):
        dim = int(rng.integers(1, 4))
        n = tuple(int(x) for x in rng.integers(3, 12, size=dim))
        ngroups = int(rng.integers(1, 4))
        groups, nlig = [], 0
        for g in range(ngroups):
            k = int(rng.integers(1, 3)) if nlig < 4 else 1
            nlig += k
            groups.append([float(rng.uniform(500, 2500)), float(rng.choice([-1, 1]) * rng.uniform(1e-4, 9e-4)),
                           [[float(rng.uniform(0.3, 1.5)), float(rng.uniform(1e-3, 2e-2)),
                             float(rng.uniform(1e-3, 2e-2)), float(rng.uniform(1e-6, 2e-5))]
                            for _ in range(k)]])
        p = dict(dim=dim, n=list(n), h=[float(rng.uniform(0.5, 2.0)) / 384 for _ in range(dim)],
                 groups=groups, s2=float(rng.uniform(1e-4, 5e-4)), rhomax=float(rng.uniform(9000, 30000)),
                 cushion=float(rng.uniform(500, 3000)), maxscale=float(rng.uniform(1, 3)),
                 cap=['tophat', 'witch'][case % 2], rhomin=1e-7, Umin=1e-7)
        ph = oracle_physics(p)
        c = OC.COracle(ph)
        u = random_state(p, 100 + case, rel=0.05)
        v = rng.standard_normal(u.size)
        f = c.dfdt(u)
        fr = O.dfdt(u, ph).reshape(-1, order='F')
        assert check_field(f, fr, ph.dof, 1e-13, cond_scale(ph, u), ncond=8.0) < 1.0, (case, p)
        shift = float(rng.uniform(1.0, 3000.0))
        c.jvp_setup(u, shift)
        assert relerr(c.jvp(v), O.jvp(u, v, shift, ph).reshape(-1, order='F'), ph.dof) < 1e-13, (case, p)
        B = O.block_diagonal(u, shift, ph)
        Mi = c.minv()
        eye = np.eye(ph.dof)
        assert max(np.abs(Mi[i] @ B[i] - eye).max() for i in range(ph.npts)) < 1e-10, (case, p)
        c.close()


def test_c_oracle_at_the_benchmark_size():
    """the grid bench.py's CPU legs run (2-D 1024^2): residual and J.v of the two oracles
    (at 256^3 the same comparison gives 0.13 of the allowed residual error and 4.9e-16 for J.v;
    run once by hand: ten seconds of numpy and some GB)"""
    p = phys84(2, (1024, 1024))
    ph = oracle_physics(p)
    c = OC.COracle(ph)
    u = random_state(p, 5)
    v = np.random.default_rng(6).standard_normal(u.size)
    f = c.dfdt(u)
    fr = O.dfdt(u, ph).reshape(-1, order='F')
    assert check_field(f, fr, ph.dof, 1e-13, cond_scale(ph, u), ncond=8.0) < 1.0
    shift = 1.0 / (O.ROSW_GAMMA * 1e-3)
    c.jvp_setup(u, shift)
    assert relerr(c.jvp(v), O.jvp(u, v, shift, ph).reshape(-1, order='F'), ph.dof) < 1e-14
    c.close()


def test_c_oracle_is_reproducible_across_thread_counts():
    """a checker must not depend on the machine: three steps (ill-conditioned state with a clamped
    point, GMRES fallback on the way) give the same bits with 1, 3 and all threads"""
    import hashlib
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    prog = (
        "import sys, hashlib; sys.path[:0] = [%r, %r]\n"
        "import numpy as np\n"
        "from helpers import phys84, oracle_physics, random_state\n"
        "from oracle import ksfd_oracle_c as OC\n"
        "p = phys84(2, (40, 28)); c = OC.COracle(oracle_physics(p))\n"
        "u = random_state(p, 9); u[0] = -1.0\n"
        "its = [c.ts_step(u, 1e-3, rtol=1e-13)[1] for _ in range(3)]\n"
        "print(OC.threads(), its, hashlib.sha256(u.tobytes()).hexdigest())\n"
    ) % (os.path.join(root, 'tests'), root)
    seen = {}
    for nt in ('1', '3', ''):
        env = dict(os.environ)
        env.pop('OMP_NUM_THREADS', None)
        if nt:
            env['OMP_NUM_THREADS'] = nt
        o = subprocess.run([sys.executable, '-c', prog], capture_output=True, text=True, timeout=300, env=env)
        assert o.returncode == 0, o.stderr[-400:]
        th, rest = o.stdout.strip().split(' ', 1)
        seen[th] = rest
    assert len(set(seen.values())) == 1, seen


@pytest.mark.parametrize('dim,n', [(1, (40,)), (2, (16, 12)), (3, (6, 5, 7))])
def test_beuler_step_vs_numpy_oracle(dim, n):
    p = phys84(dim, n)
    ph = oracle_physics(p)
    c = OC.COracle(ph)
    u = random_state(p, 2)
    un = c.beuler_step(u, 1e-3, rtol=1e-13)
    ur = O.beuler_step(u, 0.0, 1e-3, ph).reshape(-1, order='F')
    assert relerr(un, ur) < 1e-11
    inc = np.abs(ur - u).max()
    assert np.abs(un - ur).max() / inc < 1e-8
    c.close()
