"""
The spectral preconditioner's algorithm (ksfd_b200/csrc/fftpc.cuh) restated in
numpy and pinned against the oracle's assembled Jacobian (the reference's
implicitIJ, KSFD/ksfdts.py:598-640, as restated in oracle/ksfd_oracle.py):

* on a uniform state the frozen-coefficient operator A0 IS the Jacobian, so
  A0^-1 A = I to rounding — this pins the symbol of the 4th-order stencil, the
  sign conventions of the arrow block and the Schur step;
* on a random state right-preconditioned GMRES with A0^-1 converges in a handful
  of steps for dt from 1e-3 to 100 (hardware independent: the CUDA path must
  show the same counts, tests/test_gpu_parity.py).
CPU only; no product code is exercised here.
"""
import numpy as np
import pytest
import scipy.sparse.linalg as spla

from helpers import oracle_physics, phys84, random_state


def spectral_inverse(u, shift, ph):
    """numpy restatement of k_fft_means_* + k_fft_prescale + k_fft_symbol_solve +
    k_fft_postscale (+ the FFTs):  M^-1 = S_R A0^-1 S_L."""
    from oracle import ksfd_oracle as O
    ua = np.asarray(u, dtype=float).reshape(ph.Vshape, order='F').copy()
    O.groom(ua, ph)
    g_rho, g_U = O.dG_of(ua, ph)
    rho = ua[0]
    kbar, cbar, gl = (rho * g_rho).mean(), (1.0 / g_rho).mean(), [g.mean() for g in g_U]
    n = ph.Vshape[1:]
    axes = tuple(range(1, ph.dim + 1))
    lam = 0.0
    for ax in range(ph.dim):
        c = np.cos(2 * np.pi * np.arange(n[ax]) / n[ax])
        sh = [1] * ph.dim
        sh[ax] = n[ax]
        lam = lam + ((32 * c - 2 * (2 * c * c - 1) - 30) / (12 * ph.h[ax] ** 2)).reshape(sh)
    ligs = ph.ligands()

    def apply(r):
        ra = np.asarray(r, dtype=float).reshape(ph.Vshape, order='F').copy()
        ra[0] /= rho                                    # S_L: rho row divided by rho
        R = np.fft.fftn(ra, axes=axes)
        schur = shift / kbar - lam
        t = R[0].copy()
        invd = []
        for l, lig in enumerate(ligs):
            invd.append(1.0 / (shift + lig['gamma'] - lig['D'] * lam))
            bd = -gl[l] * lam * invd[l]
            schur = schur + bd * lig['s'] * cbar
            t = t - bd * R[1 + l]
        Z = np.empty_like(R)
        Z[0] = t / schur
        for l, lig in enumerate(ligs):
            Z[1 + l] = (R[1 + l] + lig['s'] * cbar * Z[0]) * invd[l]
        z = np.fft.ifftn(Z, axes=axes).real
        z[0] /= g_rho                                   # S_R: v_rho = y_0 / g_rho
        return z.reshape(-1, order='F')
    return apply


def patterned_state():
    """options84 at t = 2.7e3 on a 96^2 tile (oracle/make_pattern_state.py): rho
    between 5e2 and the density cap 2.6e4 — the pattern phase."""
    import os
    from helpers import GOLD
    g = np.load(os.path.join(GOLD, 'host_pattern96.npz'))
    return g['u'], tuple(int(x) for x in g['n'])


CASES = [('1d', phys84(1, (128,), h=1.0 / 128)), ('2d', phys84(2, (40, 32))),
         ('3d', phys84(3, (12, 10, 14)))]


@pytest.mark.parametrize('label,p', CASES)
def test_exact_on_uniform_state(label, p):
    from oracle import ksfd_oracle as O
    ph = oracle_physics(p)
    u = np.full(ph.Vshape, 9000.0).reshape(-1, order='F')
    v = np.random.default_rng(3).standard_normal(u.size)
    for dt in (1e-3, 1.0, 100.0):
        shift = 1.0 / (O.ROSW_GAMMA * dt)
        z = spectral_inverse(u, shift, ph)(O.ijacobian(u, shift, ph) @ v)
        assert np.linalg.norm(z - v) <= 1e-12 * np.linalg.norm(v), (label, dt)


@pytest.mark.parametrize('label,p', CASES)
def test_gmres_step_counts_on_random_state(label, p):
    from oracle import ksfd_oracle as O
    ph = oracle_physics(p)
    u = random_state(p, 21)
    b = np.random.default_rng(22).standard_normal(u.size)
    for dt, most in ((1e-3, 6), (1.0, 15), (100.0, 25)):
        shift = 1.0 / (O.ROSW_GAMMA * dt)
        A = O.ijacobian(u, shift, ph).tocsr()
        M = spectral_inverse(u, shift, ph)
        count = [0]

        def cb(_):
            count[0] += 1
        op = spla.LinearOperator(A.shape, matvec=lambda x: A @ M(x))
        y, info = spla.gmres(op, b, rtol=1e-10, restart=30, maxiter=10, callback=cb,
                             callback_type='pr_norm')
        assert info == 0 and count[0] <= most, (label, dt, count[0])
        x = M(y)
        assert np.linalg.norm(b - A @ x) <= 1e-9 * np.linalg.norm(b)


def test_scalings_keep_step_counts_low_on_a_patterned_state():
    """rho from 5e2 to 2.6e4: the diagonal scalings (row / rho, unknown
    g_rho*v_rho) halve the step count of the unscaled frozen-coefficient inverse
    (51 / 60 steps at dt = 5 / 20) and stay far below point-block Jacobi (which
    does not converge in 600 steps here)."""
    from oracle import ksfd_oracle as O
    u, n = patterned_state()
    p = phys84(2, n)
    ph = oracle_physics(p)
    b = O.dfdt(u, ph).reshape(-1, order='F')
    for dt, most in ((5.0, 30), (20.0, 50)):
        shift = 1.0 / (O.ROSW_GAMMA * dt)
        A = O.ijacobian(u, shift, ph).tocsr()
        M = spectral_inverse(u, shift, ph)
        count = [0]

        def cb(_):
            count[0] += 1
        op = spla.LinearOperator(A.shape, matvec=lambda x: A @ M(x))
        y, info = spla.gmres(op, b, rtol=1e-8, restart=30, maxiter=4, callback=cb,
                             callback_type='pr_norm')
        assert info == 0 and count[0] <= most, (dt, count[0])


# ---------------------------------------------------------------------------
# slab-distributed transform (fftpc.cuh: k_fft_pack, k_fft_transpose, the
# all-to-all of fftpc_apply): the index maps restated in numpy for P emulated
# ranks must reproduce the global FFT and invert exactly
# ---------------------------------------------------------------------------
def _share_start(PS, q, P):
    return PS * q // P


def _share_owner(PS, s, P):
    q = ((s + 1) * P - 1) // PS
    while q > 0 and _share_start(PS, q, P) > s:
        q -= 1
    while q + 1 < P and _share_start(PS, q + 1, P) <= s:
        q += 1
    return q


@pytest.mark.parametrize('dim,n,P', [(2, (10, 13), 3), (2, (8, 8), 4), (3, (6, 5, 7), 2),
                                     (3, (4, 6, 9), 4)])
def test_slab_transform_index_maps(dim, n, P):
    rng = np.random.default_rng(7)
    dof = 3
    NL = n[-1]
    plane = n[:-1]                                   # (nx,) or (nx, ny)
    nxh = n[0] // 2 + 1
    PS = nxh if dim == 2 else n[1] * nxh
    # global field [k][c][plane in (y, x) order]
    X = rng.standard_normal((NL, dof) + tuple(reversed(plane)))
    k0 = [r * (NL // P) + min(r, NL % P) for r in range(P + 1)]
    assert k0[P] == NL
    for s in range(PS):                              # the owner formula inverts the share starts
        q = _share_owner(PS, s, P)
        assert _share_start(PS, q, P) <= s < _share_start(PS, q + 1, P)
    # plane transforms of the own planes: A[(k*dof + c)*PS + s]
    A = []
    for r in range(P):
        loc = X[k0[r]:k0[r + 1]]
        spec = np.fft.rfft(loc, axis=-1) if dim == 2 else np.fft.rfft2(loc, axes=(-2, -1))
        A.append(spec.reshape(-1))
    # pack per destination (k_fft_pack), all-to-all, R[(k*dof + c)*nsq + s_loc] with k global
    R = []
    for q in range(P):
        s0, s1 = _share_start(PS, q, P), _share_start(PS, q + 1, P)
        nsq = s1 - s0
        Rq = np.empty(NL * dof * nsq, dtype=complex)
        for p in range(P):
            nloc = k0[p + 1] - k0[p]
            B = np.empty(nloc * dof * PS, dtype=complex)
            for e in range(nloc * dof * PS):
                kc, s = divmod(e, PS)
                d = _share_owner(PS, s, P)
                ds0 = _share_start(PS, d, P)
                dn = _share_start(PS, d + 1, P) - ds0
                B[nloc * dof * ds0 + kc * dn + (s - ds0)] = A[p][e]
            blk = B[nloc * dof * s0: nloc * dof * s0 + nloc * dof * nsq]     # what p sends to q
            Rq[dof * nsq * k0[p]: dof * nsq * k0[p] + blk.size] = blk        # where q receives it
        R.append((Rq, s0, nsq))
    # transpose (k_fft_transpose), transform along the last axis, compare with the global FFT
    G = np.fft.rfft(X, axis=-1) if dim == 2 else np.fft.rfft2(X, axes=(-2, -1))
    G = np.fft.fft(G.reshape(NL, dof, PS), axis=0)                           # [k][c][s]
    back = [None] * P
    for q, (Rq, s0, nsq) in enumerate(R):
        T = np.empty(dof * nsq * NL, dtype=complex)
        for e in range(NL * dof * nsq):
            kc, s = divmod(e, nsq)
            k, c = divmod(kc, dof)
            T[(c * nsq + s) * NL + k] = Rq[e]
        T = np.fft.fft(T.reshape(dof, nsq, NL), axis=-1)
        assert np.allclose(T, np.transpose(G[:, :, s0:s0 + nsq], (1, 2, 0)), atol=1e-10)
        T = np.fft.ifft(T, axis=-1).reshape(-1)
        Rb = np.empty_like(Rq)
        for e in range(NL * dof * nsq):
            kc, s = divmod(e, nsq)
            k, c = divmod(kc, dof)
            Rb[e] = T[(c * nsq + s) * NL + k]
        back[q] = Rb
    # the way back: share q of rank p's planes returns to rank p, unpack, inverse plane transform
    for p in range(P):
        nloc = k0[p + 1] - k0[p]
        B = np.empty(nloc * dof * PS, dtype=complex)
        for q in range(P):
            s0, nsq = R[q][1], R[q][2]
            blk = back[q][dof * nsq * k0[p]: dof * nsq * k0[p] + nloc * dof * nsq]
            B[nloc * dof * s0: nloc * dof * s0 + blk.size] = blk
        Ab = np.empty_like(B)
        for e in range(nloc * dof * PS):
            kc, s = divmod(e, PS)
            d = _share_owner(PS, s, P)
            ds0 = _share_start(PS, d, P)
            dn = _share_start(PS, d + 1, P) - ds0
            Ab[e] = B[nloc * dof * ds0 + kc * dn + (s - ds0)]
        shape = (nloc, dof) + ((nxh,) if dim == 2 else (n[1], nxh))
        loc = (np.fft.irfft(Ab.reshape(shape), n=n[0], axis=-1) if dim == 2
               else np.fft.irfft2(Ab.reshape(shape), s=(n[1], n[0]), axes=(-2, -1)))
        assert np.allclose(loc, X[k0[p]:k0[p + 1]], atol=1e-12)




@pytest.fixture(scope='module')
def share_check_exe(tmp_path_factory):
    """tests/fft_share_check.cu (host build of the __host__ __device__ functions of
    fftpc.cuh) compiled once with nvcc; no GPU needed to run it."""
    import os
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('needs nvcc')
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path_factory.mktemp('fftchk') / 'share_check')
    subprocess.run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-std=c++17',
                    '-I', os.path.join(os.path.dirname(here), 'ksfd_b200', 'csrc'),
                    os.path.join(here, 'fft_share_check.cu'), '-o', exe], check=True,
                   capture_output=True)
    return exe

def test_share_functions_as_compiled(share_check_exe):
    """fft_share_start / fft_share_owner of fftpc.cuh (host side of the __host__
    __device__ functions, built with nvcc; no GPU needed): every plane wave number
    has exactly one owner and the shares tile [0, PS)."""
    import subprocess
    exe = share_check_exe
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and 'bad 0' in out.stdout, out.stdout
    # the compiled index maps of k_fft_pack / k_fft_transpose against the formulas the
    # emulation above (test_slab_transform_index_maps) uses
    for nloc, dof, PS, P in ((5, 3, 13, 3), (4, 2, 40, 4), (7, 3, 9, 2)):
        got = [int(x) for x in subprocess.run([exe, 'pack', str(nloc), str(dof), str(PS), str(P)],
                                              capture_output=True, text=True).stdout.split()]
        want = []
        for e in range(nloc * dof * PS):
            kc, sw = divmod(e, PS)
            d = _share_owner(PS, sw, P)
            ds0 = _share_start(PS, d, P)
            want.append(nloc * dof * ds0 + kc * (_share_start(PS, d + 1, P) - ds0) + (sw - ds0))
        assert got == want and sorted(got) == list(range(nloc * dof * PS))
    for NL, dof, nsq in ((9, 3, 4), (16, 2, 7)):
        got = [int(x) for x in subprocess.run([exe, 'transpose', str(NL), str(dof), str(nsq)],
                                              capture_output=True, text=True).stdout.split()]
        want = []
        for e in range(NL * dof * nsq):
            kc, sw = divmod(e, nsq)
            k, c = divmod(kc, dof)
            want.append((c * nsq + sw) * NL + k)
        assert got == want and sorted(got) == list(range(NL * dof * nsq))


def _numpy_symbol_solve(R, shift, c2, s, gamma, D, means, n):
    """the per-wave-number solve of spectral_inverse on spectra R[c][k_last..k_0h]
    (unnormalised transforms: the 1/N is folded in, as in the kernel)"""
    dim = len(n)
    nxh = n[0] // 2 + 1
    lam = 0.0
    for ax in range(dim):
        m = nxh if ax == 0 else n[ax]
        c = np.cos(2 * np.pi * np.arange(m) / n[ax])
        sh = [1] * dim
        sh[dim - 1 - ax] = m
        lam = lam + (c2[ax] * (32 * c - 2 * (2 * c * c - 1) - 30)).reshape(sh)
    N = float(np.prod(n))
    kbar, cbar = means[0], means[1]
    nlig = R.shape[0] - 1
    schur = shift / kbar - lam
    t = R[0].copy()
    invd = []
    for l in range(nlig):
        invd.append(1.0 / (shift + gamma[l] - D[l] * lam))
        bd = -means[2 + l] * lam * invd[l]
        schur = schur + bd * s[l] * cbar
        t = t - bd * R[1 + l]
    Z = np.empty_like(R)
    Z[0] = t / schur / N
    for l in range(nlig):
        Z[1 + l] = (R[1 + l] / N + s[l] * cbar * Z[0]) * invd[l]
    return Z


@pytest.mark.parametrize('n', [(16,), (10, 7), (6, 5, 4)])
def test_symbol_solve_as_compiled(n, tmp_path, share_check_exe):
    """fft_symbol_elem of fftpc.cuh built for the host (nvcc, no GPU) against the numpy
    restatement, in the single-rank layout [c][k2][k1][k0] and in the slab-distributed
    layout [c][s_loc][k] of every rank of a 3-rank split."""
    import subprocess
    exe = share_check_exe
    rng = np.random.default_rng(11)
    dim, dof, ML = len(n), 3, 7
    nxh = n[0] // 2 + 1
    shape = tuple(reversed(n[1:])) + (nxh,)                 # [k2][k1][k0h]
    R = rng.standard_normal((dof,) + shape) + 1j * rng.standard_normal((dof,) + shape)
    shift = 0.37
    c2 = [1.0 / (12 * 0.01 ** 2)] * 3
    s = [0.01, 0.001] + [0.0] * 5
    gamma = [0.01, 0.001] + [0.0] * 5
    D = [1e-6, 1e-5] + [0.0] * 5
    means = [2.9e-4, 3.1e7, -5.2e-8, 5.1e-8] + [0.0] * 5
    want = _numpy_symbol_solve(R, shift, c2, s, gamma, D, means, n)

    def run(hd, spec):
        fin, fout = str(tmp_path / 'in.bin'), str(tmp_path / 'out.bin')
        with open(fin, 'wb') as f:
            np.array(hd, dtype=np.int32).tofile(f)
            np.array([shift] + c2 + s + gamma + D + means, dtype=np.float64).tofile(f)
            np.ascontiguousarray(spec, dtype=np.complex128).tofile(f)
        subprocess.run([exe, 'symbol', fin, fout], check=True)
        return np.fromfile(fout, dtype=np.complex128).reshape(spec.shape)

    n3 = list(n) + [1] * (3 - dim)
    got = run([dof, dof - 1] + n3 + [0, 0, 0, 0], R)
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    if dim >= 2:
        # distributed: plane index s = k0 (2-D) or k1*nxh + k0 (3-D), last axis k
        NL = n[-1]
        PS = nxh if dim == 2 else n[1] * nxh
        Rp = R.reshape(dof, NL, PS)                        # [c][k][s]
        Wp = want.reshape(dof, NL, PS)
        P = 3
        for q in range(P):
            s0, s1 = _share_start(PS, q, P), _share_start(PS, q + 1, P)
            T = np.transpose(Rp[:, :, s0:s1], (0, 2, 1))   # [c][s_loc][k]
            got = run([dof, dof - 1] + n3 + [1, s0, s1 - s0, NL], T)
            assert np.allclose(got, np.transpose(Wp[:, :, s0:s1], (0, 2, 1)), rtol=1e-12, atol=0)
