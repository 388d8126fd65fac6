"""Fourier: the (unnormalised) DFT matrix.  Mirrors fastmat/Fourier.pyx.

forward  = DFT along axis 0 (fastmat/Fourier.pyx:208-214; the reference calls numpy's pocketfft),
backward = conj(F conj(x)) = N * ifft(x) (:235-238).  Orders that do not factor into the device engine's radices
(2, 3, 4, 5, 7, 8, 11, 13, 16) are transformed with a chirp-z / Bluestein convolution (:215-231) of power-of-two
inner length; everything runs in hand-written CUDA behind ``fmb_fourier_plan_create`` / ``fmb_plan_apply``.

dtype: the reference declares complex128 (:159-161); here the matrix type is complex64 so that float32 / complex64
inputs stay in single precision (output = promote(input, complex64)), as numpy >= 2 does for bare ``np.fft.fft``.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import lib, check, FORWARD, BACKWARD
from .Matrix import Matrix, plan_apply, fft_out_type, fft_in_prepare
from .core import types as _t


class Fourier(Matrix):

    def __init__(self, order, **options):
        optimize = bool(options.get('optimize', True))
        maxStage = int(options.get('maxStage', 4))
        order = int(order)
        if order < 1:
            raise ValueError("Fourier order cannot be smaller than 1.")
        self._default_device()
        self._order = order
        h = ctypes.c_void_p()
        check(lib.fmb_fourier_plan_create(ctypes.byref(h), order, int(optimize), maxStage))
        self._plan = _lib.Plan(h)
        self._numL = int(self._plan.info.bluestein)          # the reference's decision (fastmat/Fourier.pyx:119-122)
        self._initProperties(order, order, np.complex64, **options)

    order = property(lambda self: self._order)

    def _apply(self, direction, x):
        ft_out = fft_out_type(_t.getFusedType(x.dtype), self._fusedType)
        return plan_apply(self._plan, direction, fft_in_prepare(x, ft_out, self._plan), self._order, ft_out)

    def _forward(self, x):
        return self._apply(FORWARD, x)

    def _backward(self, x):
        return self._apply(BACKWARD, x)

    # analytic overrides: fastmat/Fourier.pyx:164-194
    def _getLargestSingularValue(self):
        return float(np.sqrt(self._order))

    def _getColNorms(self):
        import torch
        return torch.full((self._order, ), float(np.sqrt(self._order)), dtype=torch.float64, device=self._default_device())

    def _getRowNorms(self):
        return self._getColNorms()

    def _getGram(self):
        from .Eye import Eye
        from .Product import Product
        return Product(Eye(self._order), float(self._order))

    def _reference(self):
        """fastmat/Fourier.pyx:241-246: exp(-2 pi i jk / N), built without any FFT."""
        import torch
        k = torch.arange(self._order, dtype=torch.float64, device=self._default_device())
        ang = torch.remainder(torch.outer(k, k), self._order) * (-2.0 * np.pi / self._order)
        return torch.complex(torch.cos(ang), torch.sin(ang))
