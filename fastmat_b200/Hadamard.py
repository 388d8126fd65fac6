"""Hadamard: natural-order (Sylvester) Walsh-Hadamard transform.  Mirrors fastmat/Hadamard.pyx.

forward = backward = unnormalised FWHT (fastmat/Hadamard.pyx:164-239) in the input's own dtype (integers wrap,
output dtype = promote(input, int8) = input).  The device kernel keeps the reference's butterfly order, so all
eight dtypes are bit-exact with the reference (``fmb_hadamard_plan_create``).
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, FORWARD
from .Matrix import Matrix, plan_apply
from .core import types as _t


class Hadamard(Matrix):

    def __init__(self, order, **options):
        order = int(order)
        if order < 1:
            raise ValueError("Hadamard: Order must be larger than 0.")
        maxOrder = 8 * 8 - 2                                     # sizeof(intsize) * 8 - 2 (fastmat/Hadamard.pyx:119)
        if order > maxOrder:
            raise ValueError("Hadamard: Order exceeds maximum for this platform: %d" % (maxOrder, ))
        self._default_device()
        self._order = order
        h = ctypes.c_void_p()
        check(lib.fmb_hadamard_plan_create(ctypes.byref(h), order))
        self._plan = _lib.Plan(h)
        self._initProperties(2 ** order, 2 ** order, np.int8, **options)
        self._forceContiguousInput = True

    order = property(lambda self: self._order)

    def _forward(self, x):
        return plan_apply(self._plan, FORWARD, x, self._numRows, _t.getFusedType(x.dtype))

    _backward = _forward

    # fastmat/Hadamard.pyx:133-156
    def _getLargestSingularValue(self):
        return float(np.sqrt(self._numRows))

    def _getColNorms(self):
        return torch.full((self._numCols, ), float(np.sqrt(self._numCols)), dtype=torch.float64, device=self._default_device())

    def _getRowNorms(self):
        return self._getColNorms()

    def _getGram(self):
        from .Eye import Eye
        from .Product import Product
        return Product(Eye(self._numRows), np.float32(self._numRows))

    def _reference(self):
        """Sylvester construction (what scipy.linalg.hadamard builds, fastmat/Hadamard.pyx:242-248)."""
        h = torch.ones((1, 1), dtype=torch.int8, device=self._default_device())
        for _ in range(self._order):
            h = torch.cat((torch.cat((h, h), dim=1), torch.cat((h, -h), dim=1)), dim=0)
        return h
