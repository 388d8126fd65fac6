"""Permutation matrix.  Mirrors fastmat/Permutation.pyx (forward x[sigma, :], backward x[tau, :], :121-125).

Exact (bit-exact) gather through the Partial index kernel.
"""
import numpy as np
import torch

from .Matrix import Matrix, plan_apply
from .Partial import _make_plan
from .Circulant import _to_host
from ._lib import FORWARD
from .core import types as _t


class Permutation(Matrix):

    def __init__(self, sigma, **options):
        sigma = np.array(_to_host(sigma))
        if sigma.ndim != 1 or not np.issubdtype(sigma.dtype, np.integer):
            raise ValueError("Not a permutation.")
        n = sigma.size
        if not np.array_equal(np.sort(sigma), np.arange(n)):
            raise ValueError("Not a permutation.")
        self._sigma = sigma.astype(np.int64)
        self._tau = np.argsort(self._sigma).astype(np.int64)
        self._default_device()
        self._fwd = _make_plan(self._sigma, n)
        self._bwd = _make_plan(self._tau, n)
        self._initProperties(n, n, np.int8, **options)

    sigma = property(lambda self: self._sigma)

    def _forward(self, x):
        return plan_apply(self._fwd, FORWARD, x, self._numRows, _t.getFusedType(x.dtype))

    def _backward(self, x):
        return plan_apply(self._bwd, FORWARD, x, self._numRows, _t.getFusedType(x.dtype))

    def _reference(self):
        return torch.eye(self._numRows, dtype=torch.int8, device=self._default_device())[torch.from_numpy(self._sigma).to(self._default_device())]
