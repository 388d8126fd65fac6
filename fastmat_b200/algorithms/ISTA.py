"""ISTA / FISTA on device tensors (fastmat/algorithms/ISTA.py:125-167, FISTA.py:131-170).

Both solvers are one proximal-gradient loop (class ISTA below); the gradient is two applies of the structured operator
through the C-ABI (``backward(forward(y) - b)``), and the update ``y - L*grad`` -> soft threshold is ONE fused CUDA kernel
(``fmb_ista_step``) instead of seven numpy sweeps.  The step size needs the largest singular
value: power iteration on the device (Matrix.largestSingularValue) replaces scipy's ARPACK ``svds``
(fastmat/Matrix.pyx:895-919).  Columns of ``arrB`` are independent problems, so a column batch shards across GPUs with no
collective (fastmat_b200.parallel).
"""
import numpy as np
import torch

from .. import _lib
from ..Matrix import Matrix, _ptr, _stream_ptr
from ..core import types as _t
from .Algorithm import Algorithm, _as_device_2d, _finish


def _flat_like(a, b):
    """True if a and b are dense with identical strides (so that element e of one is element e of the other)."""
    return a.shape == b.shape and a.stride() == b.stride() and (a.is_contiguous() or a.t().is_contiguous())


def ista_step(x, grad, numL, alpha, want_step=True):
    """(step, xnew) with step = x - numL*grad (grad None: step = x), xnew = softThreshold(step, alpha); one kernel."""
    if grad is not None and not _flat_like(x, grad):
        grad = grad.contiguous()
        x = x.contiguous()
    elif not (x.is_contiguous() or x.t().is_contiguous()):
        x = x.contiguous()
    ft = _t.getFusedType(x.dtype)
    xnew = torch.empty_like(x)
    step = torch.empty_like(x) if want_step else None
    _lib.check(_lib.lib.fmb_ista_step(_ptr(x), _ptr(grad) if grad is not None else None, _ptr(step) if step is not None else None,
                                      _ptr(xnew), x.numel(), float(numL), float(alpha), ft, _stream_ptr(x.device)))
    return step, xnew


class ISTA(Algorithm):
    """min ||Ax - b||_2^2 + lambda ||x||_1 by proximal gradient steps (what fastmat/algorithms/ISTA.py:125-167 computes).

    ISTA and FISTA are ONE loop here: a proximal gradient step from the extrapolation point ``y`` followed by the momentum
    update of ``y``; plain ISTA is the schedule with momentum 0 (``y`` is the iterate itself).  Per step: two applies of the
    operator through the C-ABI and one fused kernel."""

    PARAMETERS = {'numLambda': 0.1, 'numMaxSteps': 100, 'cbStep': None}
    _accelerated = False                     # FISTA: Nesterov's t-sequence (fastmat/algorithms/FISTA.py:131-170)

    def __init__(self, fmatA, **kwargs):
        if not isinstance(fmatA, Matrix):
            raise TypeError("fmatA must be a fastmat matrix")
        self.fmatA = fmatA
        super(ISTA, self).__init__(**kwargs)

    def softThreshold(self, arrX, numAlpha):
        """x * max(|x| - alpha, 0) / |x| (ISTA.py:113-123), one kernel."""
        return ista_step(arrX, None, 0.0, numAlpha, want_step=False)[1]

    def _work_dtype(self, b):
        # np.promote_types(np.float32, arrB.dtype) (ISTA.py:145-148), further promoted with the operator's dtype so that
        # the iterate can hold A^H(.) (the reference gets this implicitly from numpy's `-`)
        ft = _t.promoteTypes(_t.promoteTypes(_t.TYPE_FLOAT32, _t.getFusedType(b.dtype)), self.fmatA.fusedType)
        return _t.getTorchType(ft)

    def _process(self, arrB):
        self.arrB, ndim, is_np = _as_device_2d(arrB, self.fmatA)
        if self.numMaxSteps <= 0:
            raise ValueError("%s would like to do at least one step for you" % type(self).__name__)
        A = self.fmatA
        self.numL = 1.0 / (A.largestSingularValue ** 2)                       # step size 1 / sigma_max^2
        tt = self._work_dtype(self.arrB)
        b = self.arrB.to(tt)
        self.arrX = torch.zeros((self.arrB.shape[1], A.numCols), dtype=tt, device=b.device).t()   # column-major
        point = self.arrX                                                     # where the gradient is taken (y)
        self.t = 1.0
        alpha = self.numL * self.numLambda * 0.5
        for self.numStep in range(self.numMaxSteps):
            grad = A.backward(A.forward(point) - b)
            previous = self.arrX
            self.arrStep, self.arrX = ista_step(point, grad.to(tt), self.numL, alpha)
            if self._accelerated:
                t_next = (1.0 + np.sqrt(1.0 + 4.0 * self.t ** 2)) / 2.0
                point = self.arrX + ((self.t - 1.0) / t_next) * (self.arrX - previous)
                self.t = t_next
                self.arrY = point
            else:
                point = self.arrX
            self._notify(self.cbStep)
            self._notify(self.cbTrace)
        # de-biasing as in the reference: the unthresholded step values on the support (ISTA.py:163, FISTA.py:169)
        self.arrResult = torch.where(self.arrX != 0, self.arrStep, self.arrX)
        if ndim == 1:
            self.arrResult = self.arrResult.reshape(-1)
        self.arrResult = _finish(self.arrResult, is_np)
        return self.arrResult


class FISTA(ISTA):
    """ISTA with Nesterov momentum (fastmat/algorithms/FISTA.py:28-170): the same loop, accelerated schedule."""

    _accelerated = True
