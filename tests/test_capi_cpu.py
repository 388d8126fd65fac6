"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/fastmat_b200.h declares, the
host-side planner is bit-identical to the reference's, and the class layer's host logic (types, argument checks)
behaves like the reference -- all without a GPU (no compute call is made)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_data

G = golden_data()


@pytest.fixture(scope='module')
def clib():
    from fastmat_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(clib):
    header = open(os.path.join(ROOT, 'include', 'fastmat_b200.h')).read()
    declared = set(re.findall(r'\b(fmb_[a-z0-9_]+)\s*\(', header))
    declared -= {'fmb_plan', 'fmb_plan_info'}
    assert len(declared) >= 18
    raw = ctypes.CDLL(clib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), 'libfastmat_b200.so does not export %s' % name
    assert declared == set(clib.SIGNATURES), (declared ^ set(clib.SIGNATURES))
    assert clib.lib.fmb_version() >= 100


def test_planner_bit_exact_through_the_c_abi(clib):
    p = G.meta['planner']
    for ms in (2, 3, 4, 5, 7):
        got = [int(clib.lib.fmb_find_optimal_fft_size(o, ms)) for o in p['orders']]
        assert got == p['opt_%d' % ms]
    got = [int(np.float32(clib.lib.fmb_fft_complexity(o)).view(np.uint32)) for o in p['orders']]
    assert got == p['complexity_bits']


def test_no_device_fails_loudly(clib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    h = ctypes.c_void_p()
    rc = clib.lib.fmb_fourier_plan_create(ctypes.byref(h), 16, 1, 4)
    assert rc == clib.FMB_ERR_CUDA and b'no CPU fallback' in clib.lib.fmb_last_error()
    with pytest.raises(RuntimeError):
        clib.check(rc)
    import fastmat_b200 as fm
    with pytest.raises(RuntimeError):
        fm.Fourier(16)
    with pytest.raises(RuntimeError):
        fm.Hadamard(4)


def test_argument_errors_do_not_need_a_device(clib):
    h = ctypes.c_void_p()
    assert clib.lib.fmb_fourier_plan_create(ctypes.byref(h), 0, 1, 4) == clib.FMB_ERR_VALUE
    assert b'Fourier order cannot be smaller than 1' in clib.lib.fmb_last_error()
    assert clib.lib.fmb_hadamard_plan_create(ctypes.byref(h), 0) == clib.FMB_ERR_VALUE
    assert clib.lib.fmb_hadamard_plan_create(ctypes.byref(h), 63) == clib.FMB_ERR_VALUE
    idx = (ctypes.c_int64 * 2)(1, 16)
    assert clib.lib.fmb_partial_plan_create(ctypes.byref(h), idx, 2, 16) == clib.FMB_ERR_VALUE
    with pytest.raises(ValueError):
        clib.check(clib.FMB_ERR_VALUE)
    with pytest.raises(TypeError):
        clib.check(clib.FMB_ERR_TYPE)
    with pytest.raises(NotImplementedError):
        clib.check(clib.FMB_ERR_NOTIMPL)


def test_type_system_matches_reference_table():
    import torch
    from fastmat_b200.core import types as t
    names = ['int8', 'int16', 'int32', 'int64', 'float32', 'float64', 'complex64', 'complex128']
    for i, a in enumerate(names):
        assert t.getFusedType(np.dtype(a)) == i
        assert t.getFusedType(getattr(torch, a)) == i
        for j, b in enumerate(names):
            assert t.promoteTypes(a, b) == names.index(np.promote_types(a, b).name)
    assert t.safeTypeExpansion(np.int8) == t.TYPE_FLOAT32 and t.safeTypeExpansion(np.int64) == t.TYPE_FLOAT64
    assert t.safeTypeExpansion(np.complex64) == t.TYPE_COMPLEX64
    for bad in (np.float16, np.uint8, np.bool_, torch.float16, torch.bfloat16, torch.uint8):
        with pytest.raises(TypeError):
            t.getFusedType(bad)
    # the reference's recorded as-is dtypes on this numpy: Hadamard keeps every dtype
    assert all(G.meta['dtypes']['Hadamard'][n] == n for n in names)
