"""LFSRCirculant: circulant matrix of a +1/-1 linear-feedback-shift-register sequence.  Mirrors fastmat/LFSRCirculant.pyx.

forward / backward (fastmat/LFSRCirculant.pyx:316-406) = scatter the rows of x to the generator-state addresses of a zeroed
2^order buffer, fast Walsh-Hadamard transform, gather from the tap-state addresses; the forward flips the input rows
1..N-1, the backward the output rows.  On the device that is three launches through the C-ABI: the exact index
scatter / gather of ``fmb_partial_plan_create`` around the FWHT of ``fmb_hadamard_plan_create``, all in the input's own
dtype (integers wrap) and therefore bit-exact with the reference.  The register stepping itself (order, period, state
and address sequences, the constructor's ValueErrors) is host integer work done once, natively, by
``fmb_lfsr_period`` / ``fmb_lfsr_sequences``.
"""
import ctypes
import warnings

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, FORWARD, BACKWARD
from .Matrix import Matrix, plan_apply
from .Hadamard import Hadamard
from .Partial import _make_plan
from .core import types as _t


def lfsr_sequences(polynomial, start, n):
    """(generator states, tap states, +1/-1 output) over n steps: fastmat/LFSRCirculant.pyx:277-314, :343-395."""
    gen = np.empty(n, dtype=np.uint32)
    tap = np.empty(n, dtype=np.uint32)
    vec = np.empty(n, dtype=np.int8)
    check(lib.fmb_lfsr_sequences(polynomial, start, n, gen.ctypes.data_as(ctypes.c_void_p),
                                 tap.ctypes.data_as(ctypes.c_void_p), vec.ctypes.data_as(ctypes.c_void_p)))
    return gen, tap, vec


def lfsr_addresses(gen, tap):
    """Index vectors of the four exact permutations of _core (fastmat/LFSRCirculant.pyx:343-395), as int64:
    (forward scatter, backward scatter, forward gather, backward gather).  Row r of the operand meets sequence step
    k(r) = r, or 0, N-1, N-2, ..., 1 when the reference flips rows 1..N-1 (forward input, backward output: :351-361)."""
    n = gen.size
    k = np.arange(n)
    flipped = np.where(k == 0, 0, n - k)
    return (gen[flipped].astype(np.int64), gen.astype(np.int64), tap.astype(np.int64), tap[flipped].astype(np.int64))


class LFSRCirculant(Matrix):

    def __init__(self, polynomial, start, **options):
        polynomial, start = int(polynomial), int(start)
        if not (0 <= polynomial < 2 ** 32) or not (0 <= start < 2 ** 32):
            raise OverflowError("LFSRCirculant: polynomial and start must fit an unsigned 32-bit register.")
        self._polynomial = polynomial
        self._order = int(lib.fmb_lfsr_order(polynomial))
        period = int(lib.fmb_lfsr_period(polynomial, start))
        if period < 0:
            check(period)                                        # ValueError with the reference's message (:201-220)
        self._start = start & ((1 << self._order) - 1)
        self._period = period
        self._default_device()
        gen, tap, vec = lfsr_sequences(polynomial, self._start, period)
        self._states, self._vecC = gen, vec
        size = 1 << self._order
        self._content = (Hadamard(self._order), )
        # forward input P(c->g) flipped, backward input, forward output P(r->t), backward output flipped
        self._scatterFlip, self._scatter, self._gather, self._gatherFlip = (
            _make_plan(idx, size) for idx in lfsr_addresses(gen, tap))
        self._initProperties(period, period, np.int8, **options)
        self._forceContiguousInput = True

    polynomial = property(lambda self: self._polynomial)
    start = property(lambda self: self._start)
    order = property(lambda self: self._order)
    period = property(lambda self: self._period)
    vecC = property(lambda self: self._vecC)
    states = property(lambda self: self._states)

    @property
    def size(self):
        warnings.warn('size is deprecated. WIll be removed in furure releases.', FutureWarning)
        return self._order

    @property
    def taps(self):
        warnings.warn('taps is deprecated. Use polynomial.', FutureWarning)
        return self._polynomial

    def _core(self, x, scatter, gather):
        ft = _t.getFusedType(x.dtype)
        data = plan_apply(scatter, BACKWARD, x, 1 << self._order, ft)            # zero + scatter (row 0 stays zero)
        data = self._content[0].forward(data)
        return plan_apply(gather, FORWARD, data, self._numRows, ft)

    def _forward(self, x):
        return self._core(x, self._scatterFlip, self._gather)

    def _backward(self, x):
        return self._core(x, self._scatter, self._gatherFlip)

    # fastmat/LFSRCirculant.pyx:240-266
    def _roll(self, shift):
        return torch.from_numpy(np.roll(self._vecC, shift)).to(self._default_device())

    def getCol(self, idx):
        if idx < 0 or idx >= self.numCols:
            raise ValueError("Column index exceeds matrix dimensions.")
        return self._roll(idx)

    def getRow(self, idx):
        if idx < 0 or idx >= self.numRows:
            raise ValueError("Row index exceeds matrix dimensions.")
        return torch.from_numpy(self._vecC[(idx - np.arange(self._numCols)) % self._numCols]).to(self._default_device())

    def _getColNorms(self):
        return torch.full((self._numCols, ), float(np.sqrt(self._numCols)), dtype=torch.float64, device=self._default_device())

    def _getRowNorms(self):
        return self._getColNorms()

    def _reference(self):
        """fastmat/LFSRCirculant.pyx:409-437: column i is the output sequence rolled by i."""
        n = self._numRows
        i = np.arange(n)
        return torch.from_numpy(self._vecC[(i[:, None] - i[None, :]) % n]).to(self._default_device())
