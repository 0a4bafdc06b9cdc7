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
SOURCES = ['ksfd.cu']
HEADERS = ['device_common.cuh', 'naive_kernels.cuh', 'march_kernels.cuh',
           'blas1_kernels.cuh', os.path.join('..', '..', 'include', 'ksfd_b200.h')]

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3',
    '-std=c++17', '-Xcompiler', '-fPIC', '-shared',
]


def nvcc_path():
    for p in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return 'nvcc'


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS + [os.path.join('..', 'build.py')]:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return False


def build(force=False, verbose=False):
    """Compile libksfd_b200.so if missing or older than its sources."""
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + ['-o', LIB] + \
        [os.path.join(CSRC, s) for s in SOURCES] + ['-ldl']
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
        print(' '.join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
