"""
Accuracy of the table-driven fp64 log / exp / logistic used by the stencil
kernels (ksfd_b200/csrc/fastmath.cuh): the header is compiled for the HOST with
g++ and compared with long double on 4e6 points (tests/fastmath_check.cpp).
CPU only.
"""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.skipif(shutil.which('g++') is None, reason='needs g++')
def test_fastmath_accuracy(tmp_path):
    exe = str(tmp_path / 'fmcheck')
    subprocess.run(['g++', '-O2', '-std=c++17', '-ffp-contract=off',
                    '-I', os.path.join(ROOT, 'ksfd_b200', 'csrc'),
                    os.path.join(HERE, 'fastmath_check.cpp'), '-o', exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
    vals = dict(zip(out[0::2], out[1::2]))
    assert float(vals['log_ulp']) <= 1.25, vals         # beyond the 2^-58 absolute floor
    assert float(vals['log_abs_small']) <= 2.0 ** -58, vals
    assert float(vals['exp_ulp']) <= 1.5, vals
    assert float(vals['logistic_ulp']) <= 4.0, vals
    assert float(vals['rcp_ulp']) <= 0.51, vals
    assert int(vals['special']) == 0, vals


def test_tables_are_reproducible(tmp_path):
    """the committed table header is what the generator script produces"""
    import runpy
    path = os.path.join(ROOT, 'ksfd_b200', 'csrc', 'fastmath_tables.h')
    before = open(path).read()
    pytest.importorskip('mpmath')
    runpy.run_path(os.path.join(ROOT, 'scripts', 'gen_fastmath_tables.py'),
                   run_name='__main__')
    assert open(path).read() == before
