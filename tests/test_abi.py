"""
The C-ABI library builds for sm_100a, loads without a GPU, and exports every
symbol include/ksfd_b200.h declares.  No compute call is made (CPU only).
"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'ksfd_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(ksfd_[a-z0-9_]+)\s*\(', txt)))


def test_header_lists_entry_points():
    syms = header_symbols()
    assert 'ksfd_residual' in syms and 'ksfd_jvp' in syms and 'ksfd_ts_step' in syms
    assert len(syms) >= 25


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for s in header_symbols():
        assert hasattr(lib, s), s


def test_python_binding_matches_header(built_lib):
    from ksfd_b200 import _lib
    assert sorted(_lib.EXPORTS) == header_symbols()
    lib = _lib.load()
    assert lib.ksfd_abi_version() == 3
    assert _lib.launch_count() == 0


def test_struct_sizes_match_c(built_lib):
    """ctypes mirrors must have the C struct sizes (checked against sizeof
    computed from the header layout rules)."""
    from ksfd_b200 import _lib
    assert ctypes.sizeof(_lib.Physics) == 16 + 6 * 8 + 2 * 7 * 8 + 8 * 4 + 4 * 7 * 8 + 2 * 15 * 8
    assert ctypes.sizeof(_lib.KspOpts) == 3 * 8 + 6 * 4
    assert ctypes.sizeof(_lib.TsOpts) == 8 + 8 * 8 + 8 + ctypes.sizeof(_lib.KspOpts)
    assert ctypes.sizeof(_lib.TsResult) == 4 * 8 + 4 * 4 + 3 * 8 + 2 * 4


def test_no_gpu_means_loud_failure(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from ksfd_b200 import core
    with pytest.raises(core.KSFDError):
        core.Context(2, (16, 16), 3)


def test_product_does_not_import_oracle():
    """The shipped package must never route through oracle/ (CPU fallback)."""
    pkg = os.path.join(ROOT, 'ksfd_b200')
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f
