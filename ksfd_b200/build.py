"""
Build the in-tree CUDA library (sm_100a only) with nvcc.

    python -m ksfd_b200.build [--force]

The .so is written next to this file (ksfd_b200/libksfd_b200.so); it is
git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libksfd_b200.so')
# (source, extra defines, object name): the marching kernels are compiled once
# per dimension so the objects build in parallel
UNITS = [('ksfd.cu', [], 'ksfd.o')] + [
    (src, ['-DKSFD_MARCH_DIM=%d' % d], '%s_d%d.o' % (src[:-3], d))
    for src in ('march_res.cu', 'march_jvp.cu', 'march_vel.cu', 'march_sweep.cu') for d in (2, 3)]
HEADERS = ['device_common.cuh', 'naive_kernels.cuh', 'march_kernels.cuh',
           'march_launch.cuh', 'tma_march.cuh', 'tma_host.h', 'ctx.h', 'blas1_kernels.cuh', 'solver_state.cuh', 'sweep_op.cuh', 'halo_push.cuh', 'fftpc.cuh', 'fastmath.cuh',
           'fastmath_tables.h',
           os.path.join('..', '..', 'include', 'ksfd_b200.h')]
OBJDIR = os.path.join(HERE, 'build')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
    '-std=c++17', '-Xcompiler', '-fPIC',
]
# experiments: e.g. KSFD_NVCC_EXTRA="-DKSFD_TMAP_PARAM=0" python -m ksfd_b200.build --force
NVCC_FLAGS += os.environ.get('KSFD_NVCC_EXTRA', '').split()


def nvcc_path():
    for p in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return 'nvcc'


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in [u[0] for u in UNITS] + HEADERS + [os.path.join('..', 'build.py')]:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def _compile(unit, verbose):
    src, defs, obj = unit
    cmd = [nvcc_path()] + NVCC_FLAGS + defs + ['-c', os.path.join(CSRC, src),
                                               '-o', os.path.join(OBJDIR, obj)]
    if verbose:
        cmd[1:1] = ['-Xptxas', '-v']
    r = subprocess.run(cmd, capture_output=True, text=True)
    return unit, r


def build(force=False, verbose=False, out=None, extra=()):
    """Compile libksfd_b200.so if missing or older than its sources.
    out / extra: an experimental build next to it (other file name, extra nvcc flags,
    own object directory), selected at run time with KSFD_B200_LIB=<path>."""
    global OBJDIR, LIB
    if out:
        saved = (OBJDIR, LIB, list(NVCC_FLAGS))
        OBJDIR = os.path.join(HERE, 'build', os.path.splitext(os.path.basename(out))[0])
        LIB = os.path.join(HERE, out)
        NVCC_FLAGS.extend(extra)
        try:
            return _build(True, verbose)
        finally:
            OBJDIR, LIB = saved[0], saved[1]
            NVCC_FLAGS[:] = saved[2]
    return _build(force, verbose)


def _build(force, verbose):
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJDIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda u: _compile(u, verbose), UNITS))
    for unit, r in results:
        if r.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s%s' % (unit[0], r.stdout, r.stderr))
        if verbose:
            print(r.stderr)
    cmd = [nvcc_path(), '-gencode', 'arch=compute_100a,code=sm_100a', '-shared',
           '-o', LIB] + [os.path.join(OBJDIR, u[2]) for u in UNITS] + ['-ldl']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    if '--out' in sys.argv:         # python -m ksfd_b200.build --out libksfd_b200_x.so -DSOME_SWITCH=1
        i = sys.argv.index('--out')
        print(build(out=sys.argv[i + 1], extra=sys.argv[i + 2:], verbose='-v' in sys.argv))
    else:
        print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
