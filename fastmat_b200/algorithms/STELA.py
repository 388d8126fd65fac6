"""STELA (soft-thresholding with exact line search) on device tensors -- fastmat/algorithms/STELA.py:128-262.

The reference's iteration, statement by statement (including its per-column active set and its exact line search);
operator applies go through the C-ABI, the soft threshold through the fused kernel of ISTA.
"""
import torch

from ..Matrix import Matrix
from ..core import types as _t
from .Algorithm import Algorithm, _as_device_2d, _finish
from .ISTA import ista_step


class STELA(Algorithm):

    def __init__(self, fmatA, **kwargs):
        if not isinstance(fmatA, Matrix):
            raise TypeError("fmatA must be a fastmat matrix")
        self.fmatA = fmatA
        self.numLambda = 0.1
        self.numMaxSteps = 100
        self.numMaxError = 1e-6
        self.cbStep = None
        self.updateParameters(**kwargs)

    def softThreshold(self, arrX, numAlpha):
        """STELA.py:116-126."""
        return ista_step(arrX, None, 0.0, numAlpha, want_step=False)[1]

    def _process(self, arrB):
        self.arrB, ndim, is_np = _as_device_2d(arrB, self.fmatA)
        if self.numMaxSteps <= 0:
            raise ValueError("STELA would like to do at least one step for you")
        A = self.fmatA
        ft = _t.promoteTypes(_t.TYPE_FLOAT64, _t.getFusedType(self.arrB.dtype))        # STELA.py:142
        tt = _t.getTorchType(ft)
        cplx = _t.isComplex(ft)
        dev = self.arrB.device
        L = self.arrB.shape[1]
        lam = float(self.numLambda)
        self.arrGamma = torch.zeros(L, dtype=torch.float64, device=dev)
        self.arrX = torch.zeros((L, A.numCols), dtype=tt, device=dev).t()
        self.arrRes = (-self.arrB).to(tt)
        self.arrBx = torch.zeros_like(self.arrX)
        self.arrABxx = torch.zeros_like(self.arrRes)
        self.arrZ = A.backward(self.arrRes).to(tt)                                      # :165
        self.arrD = (1.0 / A.colNorms.to(torch.float64) ** 2).reshape(-1, 1)            # :168
        self.arrActive = torch.ones(L, dtype=torch.bool, device=dev)

        def finish():
            res = self.arrX.reshape(-1) if ndim == 1 else self.arrX
            return _finish(res, is_np)

        for self.numStep in range(self.numMaxSteps):
            self.arrGrad = self.arrD * self.arrX - self.arrZ                            # (17)  :181
            diff = torch.clamp(self.arrZ.real - self.arrX.real, -lam, lam)              # :184-194
            if cplx:
                diff = torch.complex(diff, torch.clamp(self.arrZ.imag - self.arrX.imag, -lam, lam))
            self.arrStop = torch.linalg.vector_norm(self.arrZ - diff, dim=0)            # :195-198
            self.arrActive = self.arrStop > self.numMaxError
            if int(self.arrActive.sum()) == 0:                                          # :203-204
                return finish()
            act = self.arrActive
            bx = self.softThreshold(self.arrGrad[:, act].contiguous(), lam) / self.arrD  # (16)  :207-212
            self.arrBx[:, act] = bx
            dx = bx - self.arrX[:, act]
            abxx = A.forward(dx).to(tt)                                                 # :215-217
            self.arrABxx[:, act] = abxx
            res_a = self.arrRes[:, act]
            num = -(torch.real(torch.sum(torch.conj(res_a) * abxx, dim=0))
                    + lam * torch.sum(bx.abs() - self.arrX[:, act].abs(), dim=0))
            gamma = torch.clamp(num / torch.sum(abxx.abs() ** 2, dim=0), 0.0, 1.0)      # (19)  :222-246
            self.arrGamma[act] = gamma
            self.arrX[:, act] = self.arrX[:, act] + dx * gamma                          # (5)   :249-251
            self.arrRes[:, act] = res_a + gamma * abxx                                  # (20)  :254-256
            self.arrZ[:, act] = A.backward(self.arrRes[:, act]).to(tt)                  # :257-259
            self.handleCallback(self.cbStep)
            self.handleCallback(self.cbTrace)
        return finish()
