"""STELA (soft-thresholding with exact line search) on device tensors - what fastmat/algorithms/STELA.py:128-262 computes.

Static shapes instead of the reference's per-step compaction: the reference gathers the still-active columns with boolean
masks before every operator apply and scatters the results back; here every step works on the full column batch and a
converged column simply gets the step length gamma = 0, which freezes its iterate, residual and gradient exactly as
skipping it does (same values, no gathers, no data-dependent shapes - one host read-back per step remains, for the
early exit).  Operator applies go through the C-ABI, the soft threshold through the fused kernel of ISTA.
"""
import torch

from ..Matrix import Matrix
from ..core import types as _t
from .Algorithm import Algorithm, _as_device_2d, _finish
from .ISTA import ista_step


class STELA(Algorithm):

    PARAMETERS = {'numLambda': 0.1, 'numMaxSteps': 100, 'numMaxError': 1e-6, 'cbStep': None}

    def __init__(self, fmatA, **kwargs):
        if not isinstance(fmatA, Matrix):
            raise TypeError("fmatA must be a fastmat matrix")
        self.fmatA = fmatA
        super(STELA, self).__init__(**kwargs)

    def softThreshold(self, arrX, numAlpha):
        return ista_step(arrX, None, 0.0, numAlpha, want_step=False)[1]

    def _process(self, arrB):
        self.arrB, ndim, is_np = _as_device_2d(arrB, self.fmatA)
        if self.numMaxSteps <= 0:
            raise ValueError("STELA would like to do at least one step for you")
        A = self.fmatA
        tt = _t.getTorchType(_t.promoteTypes(_t.TYPE_FLOAT64, _t.getFusedType(self.arrB.dtype)))
        dev = self.arrB.device
        L = self.arrB.shape[1]
        lam = float(self.numLambda)
        self.arrX = torch.zeros((L, A.numCols), dtype=tt, device=dev).t()               # column-major iterate
        self.arrRes = (-self.arrB).to(tt)                                               # A x - b
        self.arrZ = A.backward(self.arrRes).to(tt)                                      # A^H (A x - b)
        self.arrD = (1.0 / A.colNorms.to(torch.float64) ** 2).reshape(-1, 1)            # 1 / ||a_j||^2
        self.arrGamma = torch.zeros(L, dtype=torch.float64, device=dev)
        self.arrActive = torch.ones(L, dtype=torch.bool, device=dev)

        def box(v):                                                                     # projection onto [-lam, lam] per part
            if v.is_complex():
                return torch.complex(torch.clamp(v.real, -lam, lam), torch.clamp(v.imag, -lam, lam))
            return torch.clamp(v, -lam, lam)

        for self.numStep in range(self.numMaxSteps):
            # stopping measure per column: distance of the gradient to the sub-differential box (STELA.py:184-198)
            self.arrStop = torch.linalg.vector_norm(self.arrZ - box(self.arrZ - self.arrX), dim=0)
            self.arrActive = self.arrStop > self.numMaxError
            if not bool(self.arrActive.any()):
                break
            # best response of the separable approximation, search direction and its image
            self.arrGrad = self.arrD * self.arrX - self.arrZ
            self.arrBx = self.softThreshold(self.arrGrad, lam) / self.arrD
            direction = self.arrBx - self.arrX
            self.arrABxx = A.forward(direction).to(tt)
            # exact line search, clipped to [0, 1]; converged columns (and degenerate directions) take no step
            slope = torch.real(torch.sum(torch.conj(self.arrRes) * self.arrABxx, dim=0)) \
                + lam * torch.sum(self.arrBx.abs() - self.arrX.abs(), dim=0)
            curvature = torch.sum(self.arrABxx.abs() ** 2, dim=0)
            gamma = torch.clamp(-slope / torch.where(curvature > 0, curvature, torch.ones_like(curvature)), 0.0, 1.0)
            self.arrGamma = torch.where(self.arrActive & (curvature > 0), gamma, torch.zeros_like(gamma))
            self.arrX = self.arrX + direction * self.arrGamma
            self.arrRes = self.arrRes + self.arrABxx * self.arrGamma
            self.arrZ = A.backward(self.arrRes).to(tt)
            self._notify(self.cbStep)
            self._notify(self.cbTrace)
        res = self.arrX.reshape(-1) if ndim == 1 else self.arrX
        return _finish(res, is_np)
