"""Circulant: y = ifft(fft(c) * fft(x)).  Mirrors fastmat/Circulant.pyx (one level fused; multi-level composed).

The reference builds the object graph Partial(Product(F.H, Diag(fft(c)/size), F)) (fastmat/Circulant.pyx:131-137,
218-221) and pays five array sweeps per apply; here the one-level case is ONE device pipeline -- FFT -> spectrum
multiply -> inverse FFT with the embedding's zero-padding and truncation folded into the first load and the last
store (``fmb_circulant_plan_create``).  Multi-level generators (tenC.ndim > 1, :138-215) are composed from
Kron / Diag / Partial exactly as the reference does.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, FORWARD, BACKWARD
from .Matrix import Matrix, plan_apply, fft_out_type, fft_in_prepare
from .core import types as _t


def _to_host(v):
    if isinstance(v, torch.Tensor):
        return v.detach().cpu().numpy()
    return np.asarray(v)


class Circulant(Matrix):

    def __init__(self, tenC, **options):
        optimize = bool(options.get('optimize', True))
        maxStage = int(options.get('maxStage', 4))
        tenC = _to_host(tenC)
        _t.getFusedType(tenC.dtype)
        self._tenC = np.array(np.squeeze(tenC), copy=True)
        if self._tenC.ndim == 0 and tenC.ndim >= 1:
            self._tenC = self._tenC.reshape(1)          # a 1 x 1 circulant stays one-dimensional
        if tenC.ndim < 1 or self._tenC.ndim < 1:
            raise ValueError("Column-definition tensor must be at least 1D.")
        self._default_device()
        ft = _t.promoteTypes(self._tenC.dtype, _t.TYPE_COMPLEX64)
        n = int(self._tenC.size)
        self._plan = None
        self._nested = None
        if self._tenC.ndim == 1:
            c = np.ascontiguousarray(self._tenC, dtype=np.complex128)
            h = ctypes.c_void_p()
            check(lib.fmb_circulant_plan_create(ctypes.byref(h), c.ctypes.data_as(ctypes.c_void_p), n, int(optimize), maxStage))
            self._plan = _lib.Plan(h)
        else:
            self._nested = self._build_multilevel(optimize, maxStage, ft)
        self._initProperties(n, n, ft, **options)

    tenC = property(lambda self: self._tenC)
    vecC = property(lambda self: self._tenC)

    # ---- multi-level: fastmat/Circulant.pyx:138-215
    def _build_multilevel(self, optimize, maxStage, ft):
        from .Fourier import Fourier
        from .Kron import Kron
        from .Diag import Diag
        from .Product import Product
        from .Partial import Partial
        dims = np.array(self._tenC.shape)
        nopt = dims.copy()
        if optimize:
            for i, d in enumerate(dims):
                opt = int(lib.fmb_find_optimal_fft_size(int(2 * d - 1), maxStage))
                if lib.fmb_fft_complexity(opt) < lib.fmb_fft_complexity(int(d)):
                    nopt[i] = opt
        that = self._tenC.astype(np.complex128)
        for ax in range(dims.size):
            if nopt[ax] > dims[ax]:                               # _preProcSlice :244-281
                z_shape = list(that.shape)
                z_shape[ax] = int(nopt[ax] - 2 * dims[ax] + 1)
                tail = np.take(that, np.arange(1, dims[ax]), axis=ax)
                that = np.concatenate((that, np.zeros(z_shape, dtype=that.dtype), tail), axis=ax)
        total = int(np.prod(nopt))
        that = np.fft.fftn(that).reshape(total) / total
        sel = np.ones(total, dtype=bool)                           # _genArrS :283-342
        ar = np.arange(total)
        for i in range(dims.size):
            sel &= np.mod(ar, int(np.prod(nopt[i:]))) < dims[i] * int(np.prod(nopt[i + 1:]))
        KN = Kron(*[Fourier(int(d), optimize=False) for d in nopt])
        dt = np.complex64 if ft == _t.TYPE_COMPLEX64 else np.complex128
        P = Product(KN.H, Diag(that.astype(dt)), KN)
        if not np.array_equal(dims, nopt):
            idx = ar[sel]
            return Partial(P, rows=idx, cols=idx)
        return P

    def _apply(self, direction, x):
        if self._nested is not None:
            return self._nested.forward(x) if direction == FORWARD else self._nested.backward(x)
        ft_out = fft_out_type(_t.getFusedType(x.dtype), self._fusedType)
        return plan_apply(self._plan, direction, fft_in_prepare(x, ft_out, self._plan), self._numRows, ft_out)

    def _forward(self, x):
        return self._apply(FORWARD, x)

    def _backward(self, x):
        return self._apply(BACKWARD, x)

    def _getColNorms(self):
        """fastmat/Circulant.pyx:229-230."""
        return torch.full((self._numCols, ), float(np.linalg.norm(self._tenC)), dtype=torch.float64,
                          device=self._default_device())

    def _getRowNorms(self):
        return self._getColNorms()

    def _reference(self):
        """Dense circulant by index placement (fastmat/Circulant.pyx:349-425), no FFT involved."""
        dims = self._tenC.shape

        def rec(t):
            n = t.shape[0]
            i, j = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
            if t.ndim == 1:
                return t[(i - j) % n]
            blocks = [rec(t[k]) for k in range(n)]
            sub = blocks[0].shape[0]
            out = np.zeros((n * sub, n * sub), dtype=t.dtype)
            for a in range(n):
                for b in range(n):
                    out[a * sub:(a + 1) * sub, b * sub:(b + 1) * sub] = blocks[(a - b) % n]
            return out
        dt = np.complex64 if self._fusedType == _t.TYPE_COMPLEX64 else np.complex128
        return torch.from_numpy(rec(self._tenC).astype(dt)).to(self._default_device())
