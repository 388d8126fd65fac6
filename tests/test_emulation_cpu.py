"""CPU tests of the KERNEL LOGIC: tests/emul/libfmb_emul.so is the same CUDA source compiled with -DFMB_EMULATE, where
every launch runs the __host__ __device__ kernel body on host threads with a real barrier (fastmat_b200/csrc/emulate.cpp).
It exists because the build container has no GPU; it is never shipped and the product library contains no such path.
Checked against the numpy oracle (which is itself pinned to the real reference by tests/test_oracle.py)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import ROOT, relerr
from oracle import fastmat_oracle as orc

EMUL = os.path.join(ROOT, 'tests', 'emul', 'libfmb_emul.so')
DT = {np.dtype(n): i for i, n in enumerate(['int8', 'int16', 'int32', 'int64', 'float32', 'float64', 'complex64', 'complex128'])}
i64 = ctypes.c_int64


@pytest.fixture(scope='module')
def lib():
    src_dir = os.path.join(ROOT, 'fastmat_b200', 'csrc')
    newest = max(os.path.getmtime(os.path.join(src_dir, f)) for f in os.listdir(src_dir))
    newest = max(newest, os.path.getmtime(os.path.join(ROOT, 'include', 'fastmat_b200.h')))
    if not os.path.exists(EMUL) or os.path.getmtime(EMUL) < newest:          # missing or older than the sources
        if shutil.which('nvcc') is None and not os.path.exists('/usr/local/cuda/bin/nvcc'):
            pytest.skip('no nvcc to build the emulation library')
        subprocess.check_call(['bash', os.path.join(ROOT, 'build.sh')], env=dict(os.environ, EMUL='1'), cwd=ROOT)
    from fastmat_b200 import _lib
    return _lib.bind(ctypes.CDLL(EMUL))


def chk(lib, rc):
    assert rc == 0, lib.fmb_last_error()


def apply(lib, plan, direction, x, out_rows, out_dtype):
    M = x.shape[1]
    y = np.zeros((out_rows, M), dtype=out_dtype, order='F' if x.flags.f_contiguous else 'C')
    wsb = lib.fmb_plan_workspace_bytes(plan, direction, M, DT[x.dtype], DT[y.dtype])
    ws = np.zeros(max(wsb, 16), dtype=np.uint8)
    chk(lib, lib.fmb_plan_apply(plan, direction, x.ctypes.data, x.strides[0] // x.itemsize, x.strides[1] // x.itemsize,
                                y.ctypes.data, y.strides[0] // y.itemsize, y.strides[1] // y.itemsize, M, DT[x.dtype],
                                DT[y.dtype], ws.ctypes.data, wsb, None))
    return y


def crand(rng, *s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


@pytest.mark.parametrize('n', [1, 2, 3, 8, 35, 64, 127, 243, 1000, 1021, 4096, 10000, 16384])
@pytest.mark.parametrize('order', ['F', 'C'])
def test_fourier_emulated(lib, n, order):
    rng = np.random.default_rng(n)
    p = ctypes.c_void_p()
    chk(lib, lib.fmb_fourier_plan_create(ctypes.byref(p), n, 1, 4))
    for dt, tol in ((np.complex128, 1e-12), (np.complex64, 5e-6)):   # the oracle (reference Bluestein) is itself ~1e-14
        x = np.asarray(crand(rng, n, 3), dtype=dt, order=order)
        assert relerr(apply(lib, p, 0, x, n, dt), orc.fourier_forward(x)) < tol
        assert relerr(apply(lib, p, 1, x, n, dt), orc.fourier_backward(x)) < tol
    lib.fmb_plan_destroy(p)


@pytest.mark.parametrize('n', [1, 7, 41, 100, 1021, 5000, 20000])
def test_circulant_emulated(lib, n):
    rng = np.random.default_rng(n)
    c = np.ascontiguousarray(crand(rng, n))
    p = ctypes.c_void_p()
    chk(lib, lib.fmb_circulant_plan_create(ctypes.byref(p), c.ctypes.data, n, 1, 4))
    for order in 'FC':
        x = np.asarray(crand(rng, n, 3), order=order)
        assert relerr(apply(lib, p, 0, x, n, np.complex128), orc.circulant_forward(c, x)) < 1e-13
        assert relerr(apply(lib, p, 1, x, n, np.complex128), orc.circulant_backward(c, x)) < 1e-13
    lib.fmb_plan_destroy(p)


@pytest.mark.parametrize('n,m', [(4, 3), (4, 41), (1, 5), (5, 1), (100, 100), (1000, 24), (9000, 9000)])
def test_toeplitz_emulated(lib, n, m):
    rng = np.random.default_rng(n * 7 + m)
    vc, vr = np.ascontiguousarray(crand(rng, n)), np.ascontiguousarray(crand(rng, m - 1))
    p = ctypes.c_void_p()
    chk(lib, lib.fmb_toeplitz_plan_create(ctypes.byref(p), vc.ctypes.data, n, vr.ctypes.data if m > 1 else None, m - 1, 1, 4))
    for order in 'FC':
        x = np.asarray(crand(rng, m, 3), order=order)
        y = np.asarray(crand(rng, n, 3), order=order)
        assert relerr(apply(lib, p, 0, x, n, np.complex128), orc.toeplitz_forward(vc, vr, x)) < 1e-13
        assert relerr(apply(lib, p, 1, y, m, np.complex128), orc.toeplitz_backward(vc, vr, y)) < 1e-13
    lib.fmb_plan_destroy(p)


@pytest.mark.parametrize('dims', [(4, 8), (5, 7), (128, 64), (100, 30)])
def test_kron_fourier_emulated(lib, dims):
    rng = np.random.default_rng(dims[0])
    p = ctypes.c_void_p()
    chk(lib, lib.fmb_kron_fourier_plan_create(ctypes.byref(p), (ctypes.c_int64 * 2)(*dims), 2))
    n = dims[0] * dims[1]
    for order in 'FC':
        x = np.asarray(crand(rng, n, 3), order=order)
        assert relerr(apply(lib, p, 0, x, n, np.complex128), orc.kron_fourier_forward(dims, x)) < 1e-13
        assert relerr(apply(lib, p, 1, x, n, np.complex128), orc.kron_fourier_backward(dims, x)) < 1e-13
    lib.fmb_plan_destroy(p)


@pytest.mark.parametrize('order', [1, 2, 6, 10, 13, 15])
def test_hadamard_emulated_bit_exact(lib, order):
    rng = np.random.default_rng(order)
    p = ctypes.c_void_p()
    chk(lib, lib.fmb_hadamard_plan_create(ctypes.byref(p), order))
    n = 1 << order
    for dt in ('int8', 'int16', 'int32', 'int64', 'float32', 'float64', 'complex64', 'complex128'):
        if dt.startswith('int'):
            x = rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, size=(n, 3), dtype=np.int64, endpoint=True).astype(dt)
        elif dt.startswith('float'):
            x = rng.standard_normal((n, 3)).astype(dt)
        else:
            x = crand(rng, n, 3).astype(dt)
        for lay in 'FC':
            xx = np.asarray(x, order=lay)
            y = apply(lib, p, 0, xx, n, xx.dtype)
            ref = orc.hadamard_forward(xx)
            assert np.array_equal(np.ascontiguousarray(y).view(np.uint8), np.ascontiguousarray(ref).view(np.uint8)), (dt, lay)
    lib.fmb_plan_destroy(p)
