"""Diag: diagonal matrix.  Mirrors fastmat/Diag.pyx (forward :149-157, backward with conj(d) :159-167)."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, FORWARD, BACKWARD
from .Matrix import Matrix, plan_apply
from .Circulant import _to_host
from .core import types as _t


class Diag(Matrix):

    def __init__(self, vecD, **options):
        vecD = _to_host(vecD)
        ft = _t.getFusedType(vecD.dtype)
        self._vecD = np.array(np.squeeze(vecD), copy=True) if vecD.size != 1 else np.array(vecD, copy=True).reshape(vecD.shape[:1] or (1, ))
        if vecD.ndim != 1 or self._vecD.ndim != 1:
            raise ValueError("Diag: Definition vector must have exactly one dimension.")
        self._default_device()
        d = np.ascontiguousarray(self._vecD)
        h = ctypes.c_void_p()
        check(lib.fmb_diag_plan_create(ctypes.byref(h), d.ctypes.data_as(ctypes.c_void_p), ft, int(d.size)))
        self._plan = _lib.Plan(h)
        self._initProperties(d.size, d.size, self._vecD.dtype, **options)
        self._forceContiguousInput = True

    vecD = property(lambda self: self._vecD)

    def _apply(self, direction, x):
        ft_out = _t.promoteTypes(x.dtype, self._fusedType)
        return plan_apply(self._plan, direction, x, self._numRows, ft_out)

    def _forward(self, x):
        return self._apply(FORWARD, x)

    def _backward(self, x):
        return self._apply(BACKWARD, x)

    def _getLargestSingularValue(self):
        return float(np.abs(self._vecD).max())

    def _getColNorms(self):
        return torch.from_numpy(np.abs(self._vecD).astype(np.float64)).to(self._default_device())

    def _getRowNorms(self):
        return self._getColNorms()

    def _getGram(self):
        return Diag(np.abs(self._vecD) ** 2)

    def _getT(self):
        return self

    def _reference(self):
        return torch.from_numpy(np.diag(self._vecD)).to(self._default_device())
