"""CPU tier: the C-ABI shared library builds, loads and exports exactly what include/lssurf_b200.h
declares; without a GPU the compute path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from lssurf_b200 import _cabi
from tests.util import ROOT


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'lssurf_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(lss_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f'{n} declared in the header but not exported'


def test_binding_table_covers_header():
    declared = set(header_functions()) - {'lss_last_error'}
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)


def test_version_and_error_string():
    lib = _cabi.load()
    assert lib.lss_version() >= 100
    assert isinstance(lib.lss_last_error(), bytes)


def test_argument_validation_without_gpu():
    """Pure host-side entry points work everywhere; bad arguments give an error code and a message."""
    lib = _cabi.load()
    g = _cabi.grid_struct([5, 7], [7, 1], [0., 0.], [4., 6.], [1., 1.], 0)
    import numpy as np
    offs = np.array([[0, 0, 0], [-1, 0, 1]], dtype=np.int32)
    n_eq = ctypes.c_int64(0)
    rc = lib.lss_stencil_count(ctypes.byref(g), 3, offs.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 1,
                               ctypes.byref(n_eq), None, None)
    assert rc == 0 and n_eq.value == 5 * 5
    g.ndim = 7
    rc = lib.lss_stencil_count(ctypes.byref(g), 3, offs.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 1,
                               ctypes.byref(n_eq), None, None)
    assert rc == -1 and b'ndim' in lib.lss_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback():
    import numpy as np
    from lssurf_b200 import RDE, fd_grid, lin_op
    assert _cabi.load().lss_device_count() == 0
    with pytest.raises(_cabi.LssError):
        _cabi.lib()
    with pytest.raises((_cabi.LssError, RuntimeError, AssertionError)):
        RDE(np.arange(10.))
    g = fd_grid([[0, 10], [0, 10]], [1, 1], name='z0')
    with pytest.raises((_cabi.LssError, RuntimeError, AssertionError)):
        lin_op(g).interp_mtx([np.array([1.5]), np.array([2.5])])


def test_product_does_not_import_oracle():
    """The product package must never reach into oracle/ (it is the checker, not a fallback)."""
    pkg = os.path.join(ROOT, 'lssurf_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, flags=re.M), f
