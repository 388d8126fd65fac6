"""Toeplitz: T[i, j] = t[(i - j) mod (n + m - 1)].  Mirrors fastmat/Toeplitz.pyx (one level fused; multi-level composed).

``Toeplitz(vecC, vecR)``: vecC is the first column (n), vecR the first row without element (0, 0), stored reversed
as in the reference (fastmat/Toeplitz.pyx:730-732).  The reference embeds into an L x L circulant and selects the
n x m corner with boolean masks through Partial (:253-316); here the zero-padding of x (m -> L) happens in the
kernel's loads and the truncation (L -> n) in its stores (``fmb_toeplitz_plan_create``).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, FORWARD, BACKWARD
from .Matrix import Matrix, plan_apply, fft_out_type, fft_in_prepare
from .Circulant import _to_host
from .core import types as _t


class Toeplitz(Matrix):

    def __init__(self, *args, **options):
        split = options.pop('split', None)
        arrSplit = np.array([] if split is None else split)
        maxStage = int(options.get('maxStage', 4))
        optimize = bool(options.get('optimize', True))
        self._plan = None
        self._nested = None
        if len(args) == 1:
            self._tenT = np.array(_to_host(args[0]), copy=True)
            _t.getFusedType(self._tenT.dtype)
        elif len(args) == 2:
            if not all(isinstance(a, (np.ndarray, torch.Tensor)) for a in args):
                raise ValueError("You must specify two 1D-ndarrays containing the column- and row-definition vectors "
                                 "or one ndarray tensor")
            if arrSplit.size != 0:
                raise ValueError("You must not define split points when supplying column- and row-definition vectors.")
            vc, vr = _to_host(args[0]), _to_host(args[1])
            dataType = np.promote_types(vc.dtype, vr.dtype)
            _t.getFusedType(dataType)
            vc = np.squeeze(vc.astype(dataType)) if vc.size != 1 else vc.astype(dataType).reshape(1)
            vr = np.squeeze(vr.astype(dataType)) if vr.size != 1 else vr.astype(dataType).reshape(vr.size)
            if vc.ndim != 1 or vr.ndim != 1:
                raise ValueError("Column- and row-definition vectors must be 1D.")
            arrSplit = np.array(vc.size)
            self._tenT = np.hstack((vc, vr))
        else:
            raise ValueError("Invalid number of arguments to Toeplitz: Expecting exactly one or two fixed arguments")
        arrDim = np.array(self._tenT.shape)
        arrSplit = np.atleast_1d(arrSplit)
        if arrSplit.size == 0:
            if not all(((ll + 1) % 2 == 0) for ll in arrDim):
                raise ValueError("Defining a tensor with non-square levels requires explicit split points.")
            arrSplit = (arrDim + 1) // 2
        if arrSplit.size != self._tenT.ndim:
            raise ValueError("The split point vector must have one entry for each dimension of the defining tensor")
        elif arrSplit.ndim != 1:
            raise ValueError("The split point vector must be 1D")
        elif any(ll < 1 or ll > arrDim[ii] for ii, ll in enumerate(arrSplit)):
            raise ValueError("Entry in split vector outside of defining tensor bounds")
        self._arrDimRows = arrSplit.astype(np.int64)
        self._arrDimCols = (arrDim - self._arrDimRows + 1).astype(np.int64)
        self._default_device()
        ft = _t.promoteTypes(self._tenT.dtype, _t.TYPE_COMPLEX64)
        n, m = int(np.prod(self._arrDimRows)), int(np.prod(self._arrDimCols))
        if self._tenT.ndim == 1:
            vc = np.ascontiguousarray(self._tenT[:n], dtype=np.complex128)
            vr = np.ascontiguousarray(self._tenT[n:], dtype=np.complex128)
            h = ctypes.c_void_p()
            check(lib.fmb_toeplitz_plan_create(ctypes.byref(h), vc.ctypes.data_as(ctypes.c_void_p), n,
                                               vr.ctypes.data_as(ctypes.c_void_p) if vr.size else None, int(vr.size),
                                               int(optimize), maxStage))
            self._plan = _lib.Plan(h)
        else:
            self._nested = self._build_multilevel(optimize, maxStage, ft)
        self._initProperties(n, m, ft, **options)

    tenT = property(lambda self: self._tenT)

    @property
    def vecC(self):
        return self._tenT[:self._arrDimRows[0]] if self._tenT.ndim == 1 else None

    @property
    def vecR(self):
        return self._tenT[self._arrDimRows[0]:] if self._tenT.ndim == 1 else None

    # ---- multi-level: fastmat/Toeplitz.pyx:214-316
    def _build_multilevel(self, optimize, maxStage, ft):
        from .Fourier import Fourier
        from .Kron import Kron
        from .Diag import Diag
        from .Product import Product
        from .Partial import Partial
        dims = np.array(self._tenT.shape)
        dopt = dims.copy()
        if optimize:
            for i, d in enumerate(dims):
                opt = int(lib.fmb_find_optimal_fft_size(int(d), maxStage))
                if lib.fmb_fft_complexity(opt) < lib.fmb_fft_complexity(int(d)):
                    dopt[i] = opt
        that = self._tenT.astype(np.complex128)
        for ax in range(dims.size):
            if dopt[ax] > dims[ax]:                                # _preProcSlice :331-369: zeros at the split point
                bp = int(self._arrDimRows[ax])
                head = np.take(that, np.arange(0, bp), axis=ax)
                tail = np.take(that, np.arange(bp, dims[ax]), axis=ax)
                z_shape = list(that.shape)
                z_shape[ax] = int(dopt[ax] - dims[ax])
                that = np.concatenate((head, np.zeros(z_shape, dtype=that.dtype), tail), axis=ax)
        total = int(np.prod(dopt))
        that = np.fft.fftn(that).reshape(total)
        F = Kron(*[Fourier(int(d), optimize=False) for d in dopt])
        dt = np.complex64 if ft == _t.TYPE_COMPLEX64 else np.complex128
        P = Product(F.H, Diag((that / total).astype(dt)), F)
        ar = np.arange(total)
        rows = ar >= 0
        cols = ar >= 0
        for i in range(dims.size):                                 # :284-313
            below = int(np.prod(dopt[i + 1:]))
            mod = np.mod(ar, int(dopt[i]) * below)
            rows &= mod < int(self._arrDimRows[i]) * below
            cols &= mod < int(self._arrDimCols[i]) * below
        return Partial(P, rows=rows, cols=cols)

    # ---- norms in closed form (what fastmat/Toeplitz.pyx:371-625 obtains by its level recursion): column j of one level
    # holds the generator entries with offsets -j .. nr-1-j, so its squared norm is a difference of cumulative sums of
    # |t|^2; for several levels the same sum runs over the squared norms of the sub-blocks.  O(size of the generator) on the
    # host instead of numCols / chunk forward applies of one-hot slabs (16384 of them at 2^19 x 2^19).
    def _norms2(self, t, level, rows):
        nr, nc = int(self._arrDimRows[level]), int(self._arrDimCols[level])
        d = nr + nc - 1
        if t.ndim == 1:
            P = np.abs(t.astype(np.complex128)) ** 2
        else:
            P = np.stack([self._norms2(t[k], level + 1, rows) for k in range(d)])
        if rows:
            nr, nc = nc, nr                                        # row i of T = column i of T^T: generator index reflected
            P = np.concatenate((P[:1], P[:0:-1]), axis=0)
        tail = P.shape[1:]
        zero = np.zeros((1, ) + tail)
        down = np.concatenate((zero, np.cumsum(P[:nr], axis=0)), axis=0)            # down[k] = sum of P[0 .. k-1]
        up = np.concatenate((zero, np.cumsum(P[:nr - 1:-1] if nc > 1 else P[:0], axis=0)), axis=0)   # up[j] = sum_{q=1..j} P[d-q]
        j = np.arange(nc)
        # offsets 0 .. nr-1-j come from the column part, offsets -q with max(1, j-nr+1) <= q <= j from the row part
        return down[np.clip(nr - j, 0, nr)] + up[j] - up[np.clip(j - nr, 0, None)]

    def _getColNorms(self):
        n2 = self._norms2(self._tenT, 0, False).reshape(-1)
        return torch.from_numpy(np.sqrt(n2)).to(self._default_device())

    def _getRowNorms(self):
        n2 = self._norms2(self._tenT, 0, True).reshape(-1)
        return torch.from_numpy(np.sqrt(n2)).to(self._default_device())

    def _apply(self, direction, x):
        if self._nested is not None:
            return self._nested.forward(x) if direction == FORWARD else self._nested.backward(x)
        ft_out = fft_out_type(_t.getFusedType(x.dtype), self._fusedType)
        rows_out = self._numRows if direction == FORWARD else self._numCols
        return plan_apply(self._plan, direction, fft_in_prepare(x, ft_out, self._plan), rows_out, ft_out)

    def _forward(self, x):
        return self._apply(FORWARD, x)

    def _backward(self, x):
        return self._apply(BACKWARD, x)

    def _reference(self):
        """fastmat/Toeplitz.pyx:628-734 by index placement: T[i, j] = t[(i - j) mod (n + m - 1)] per level."""
        def rec(t, level):
            nr, nc = int(self._arrDimRows[level]), int(self._arrDimCols[level])
            d = nr + nc - 1
            i, j = np.meshgrid(np.arange(nr), np.arange(nc), indexing='ij')
            if t.ndim == 1:
                return t[(i - j) % d]
            blocks = [rec(t[k], level + 1) for k in range(d)]
            br, bc = blocks[0].shape
            out = np.zeros((nr * br, nc * bc), dtype=t.dtype)
            for a in range(nr):
                for b in range(nc):
                    out[a * br:(a + 1) * br, b * bc:(b + 1) * bc] = blocks[(a - b) % d]
            return out
        dt = np.complex64 if self._fusedType == _t.TYPE_COMPLEX64 else np.complex128
        return torch.from_numpy(rec(self._tenT, 0).astype(dt)).to(self._default_device())
