"""Kronecker product of square matrices.  Mirrors fastmat/Kron.pyx.

forward (fastmat/Kron.pyx:267-303) is a chain of mode-i products on the row-major index (i1, ..., ik).  When both
factors of a two-factor product are Fourier matrices the chain is the 2-D DFT of the row-major reshaped column and
runs as ONE two-pass device pipeline (``fmb_kron_fourier_plan_create``); every other combination applies the factors
one mode at a time through their own device transforms.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, FORWARD, BACKWARD
from .Matrix import Matrix, plan_apply, cast, fft_in_prepare
from .core import types as _t


class Kron(Matrix):

    def __init__(self, *matrices, **options):
        self._content = tuple(matrices)
        if len(matrices) < 2:
            raise ValueError("Kronecker: Product must have at least two terms")
        ft = _t.TYPE_INT8
        numRows = 1
        for f in self._content:
            if not isinstance(f, Matrix):
                raise TypeError("Kronecker: Term is not a Matrix.")
            if f.numRows != f.numCols:
                raise ValueError("Kronecker: Product terms must be symmetric")
            numRows *= f.numRows
            ft = _t.promoteTypes(ft, f.fusedType)
        self._dims = tuple(f.numRows for f in self._content)
        expansion = options.get('typeExpansion', _t.safeTypeExpansion(ft))
        if expansion is not None:
            ft = _t.promoteTypes(ft, expansion)
        self._plan = None
        from .Fourier import Fourier
        if len(self._content) == 2 and all(type(f) is Fourier for f in self._content):
            dims = (ctypes.c_int64 * 2)(*self._dims)
            h = ctypes.c_void_p()
            rc = lib.fmb_kron_fourier_plan_create(ctypes.byref(h), dims, 2)
            if rc == 0:
                self._plan = _lib.Plan(h)
            elif rc != _lib.FMB_ERR_NOTIMPL:
                check(rc)
        self._initProperties(numRows, numRows, ft, **options)
        self._widenInputDatatype = True                      # fastmat/Kron.pyx:129

    def _chain(self, x, backward):
        M = x.shape[1]
        n_total = self._numRows
        # tensor view (d1, ..., dk, M) of the row-major index; one mode product per factor (Kron.pyx:279-300)
        data = x.reshape(self._dims + (M, ))
        for ax, term in enumerate(self._content):
            moved = data.movedim(ax, 0)
            shp = moved.shape
            flat = moved.reshape(shp[0], -1)
            out = term.backward(flat) if backward else term.forward(flat)
            data = out.reshape((shp[0], ) + tuple(shp[1:])).movedim(0, ax)
        return data.reshape(n_total, M)

    def _forward(self, x):
        if self._plan is not None:
            ft_out = _t.promoteTypes(x.dtype, _t.TYPE_COMPLEX64)
            return plan_apply(self._plan, FORWARD, fft_in_prepare(x, ft_out, self._plan), self._numRows, ft_out)
        return self._chain(x, False)

    def _backward(self, x):
        if self._plan is not None:
            ft_out = _t.promoteTypes(x.dtype, _t.TYPE_COMPLEX64)
            return plan_apply(self._plan, BACKWARD, fft_in_prepare(x, ft_out, self._plan), self._numRows, ft_out)
        return self._chain(x, True)

    # analytic overrides: fastmat/Kron.pyx:155-183
    def _getLargestSingularValue(self):
        v = 1.0
        for f in self._content:
            v *= float(f.largestSingularValue)
        return v

    def _getColNorms(self):
        n = self._content[0].colNorms
        for f in self._content[1:]:
            n = torch.kron(n, f.colNorms)
        return n

    def _getRowNorms(self):
        n = self._content[0].rowNorms
        for f in self._content[1:]:
            n = torch.kron(n, f.rowNorms)
        return n

    def _reference(self):
        """fastmat/Kron.pyx:344-353: np.kron of the factor references."""
        arr = None
        for f in self._content:
            r = f.reference()
            arr = r if arr is None else torch.kron(arr.to(_t.promoteTorch(arr.dtype, r.dtype)),
                                                   r.to(_t.promoteTorch(arr.dtype, r.dtype)))
        return arr
