"""Orthogonal Matching Pursuit on device tensors (fastmat/algorithms/OMP.pyx:122-253).

Same iteration as the reference, batched over the L right-hand sides: correlation ``|C^H r|`` with the column-normalised
operator (one backward apply through the C-ABI), arg-max per column, fetch of the picked atoms (one forward apply of
one-hot columns), and the rank-one update of the pseudo inverse restricted to the support (einsum on the device).  The
work type is the reference's: ``promote(promote(C.dtype, b.dtype), float64)``.
"""
import numpy as np
import torch

from ..Matrix import Matrix
from ..core import types as _t
from .Algorithm import Algorithm, _as_device_2d, _finish


class OMP(Algorithm):

    def __init__(self, fmatA, **kwargs):
        if not isinstance(fmatA, Matrix):
            raise TypeError("fmatA must be a fastmat matrix")
        self.fmatA = fmatA
        self.numMaxSteps = 0
        self.cbStep = None
        self.updateParameters(**kwargs)

    def _process(self, arrB):
        self.arrB, ndim, is_np = _as_device_2d(arrB, self.fmatA)
        if self.numMaxSteps <= 0:
            raise ValueError("OMP would like to do at least one step for you")
        A = self.fmatA
        K = int(self.numMaxSteps)
        self.numN, self.numM, self.numL = A.numRows, A.numCols, self.arrB.shape[1]
        dev = self.arrB.device
        self.fmatC = A.colNormalized                                          # OMP.pyx:149
        ft = _t.promoteTypes(_t.promoteTypes(self.fmatC.fusedType, _t.getFusedType(self.arrB.dtype)), _t.TYPE_FLOAT64)
        tt = _t.getTorchType(ft)                                               # OMP.pyx:152-155
        self.returnType = _t.getNumpyType(ft)
        N, M, L = self.numN, self.numM, self.numL
        b = self.arrB.to(tt)
        self.arrXtmp = torch.zeros((K, L), dtype=tt, device=dev)
        self.arrResidual = b.clone()
        self.arrSupport = torch.empty((K, L), dtype=torch.long, device=dev)
        self.matPinv = torch.zeros((K, N, L), dtype=tt, device=dev)
        self.arrA = torch.zeros((N, K, L), dtype=tt, device=dev)
        cols = torch.arange(L, device=dev)
        for self.numStep in range(K):
            ii = self.numStep
            self.arrC = self.fmatC.backward(self.arrResidual).abs()           # OMP.pyx:196
            self.newIndex = torch.argmax(self.arrC, dim=0)                    # :199 (first maximum, like np.argmax)
            self.arrSupport[ii, :] = self.newIndex
            self.newCols = A.getCols(self.newIndex).to(tt)                    # :205
            self.arrA[:, ii, :] = self.newCols
            if ii == 0:                                                       # :211-218
                self.v2 = self.newCols
                self.v2n = (self.v2 / torch.linalg.vector_norm(self.v2, dim=0) ** 2).conj()
                self.v2y = torch.einsum('ji,ji->i', self.v2n, b)
                self.arrXtmp[0, :] = self.v2y
                self.matPinv[0, :, :] = self.v2n
            else:                                                             # :219-238
                self.v1 = torch.einsum('ijk,jk->ik', self.matPinv[:ii], self.newCols)
                self.v2 = self.newCols - torch.einsum('ijk,jk->ik', self.arrA[:, :ii, :], self.v1)
                self.v2n = (self.v2 / torch.linalg.vector_norm(self.v2, dim=0) ** 2).conj()
                self.v2y = torch.einsum('ji,ji->i', self.v2n, b)
                self.arrXtmp[:ii, :] -= self.v2y * self.v1
                self.arrXtmp[ii, :] += self.v2y
                self.matPinv[:ii] -= torch.einsum('ik,jk->jik', self.v2n, self.v1)
                self.matPinv[ii] = self.v2n
            self.arrResidual = self.arrResidual - self.v2y * self.v2          # :241
            self.handleCallback(self.cbStep)
            self.handleCallback(self.cbTrace)
        self.arrX = torch.zeros((M, L), dtype=tt, device=dev)
        self.arrX[self.arrSupport, cols] = self.arrXtmp                       # :249-250
        res = self.arrX.reshape(-1) if ndim == 1 else self.arrX
        return _finish(res, is_np)
