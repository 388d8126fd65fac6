"""Orthogonal Matching Pursuit on device tensors, batched over the right-hand sides.

What it computes is the reference's OMP (fastmat/algorithms/OMP.pyx:122-253): per step the atom of the column-normalised
operator that correlates best with the residual joins the support (:196-199), and the coefficients on the support are the
least-squares solution (:211-241).  How it computes it is different by design:

* selection: one backward apply of the normalised operator through the C-ABI, then ONE kernel (``fmb_abs_argmax``) that
  reduces |.| and arg-max per column in a single sweep - the magnitude array the reference builds is never formed;
* least squares: instead of the reference's explicit pseudo-inverse with rank-one updates (two K x N x L arrays, three
  sweeps over them per step) the support atoms are orthonormalised incrementally (Gram-Schmidt with one
  re-orthogonalisation: Q holds an orthonormal basis, R its triangular factor), the residual is updated with the new basis
  vector only, and the K coefficients come from ONE batched triangular solve ``R x = Q^H b`` at the end.  One N x K x L
  array, swept by two memory-bound kernels of the library per orthogonalisation (``fmb_gs_project`` / ``fmb_gs_subtract``)
  instead of four einsums over two such arrays.  Same minimiser, so the results agree with the reference's to rounding
  (tests/test_gpu_algorithms.py: identical support, values to 1e-9).

Columns of ``arrB`` are independent problems: a batch shards across GPUs without a collective (fastmat_b200.parallel).
The work type is the reference's: ``promote(promote(C.dtype, b.dtype), float64)``.
"""
import torch

from .. import _lib
from ..Matrix import Matrix, _ptr, _stream_ptr
from ..core import types as _t
from .Algorithm import Algorithm, _as_device_2d, _finish


def abs_argmax(x):
    """Row index of the largest magnitude of every column of a column-major 2-D CUDA tensor (first one on ties)."""
    if x.ndim != 2:
        raise ValueError("abs_argmax expects a 2-D tensor")
    if x.stride(0) != 1 and x.shape[0] > 1:
        x = x.t().contiguous().t()
    rows, cols = x.shape
    out = torch.empty(cols, dtype=torch.int64, device=x.device)
    ws_bytes = int(_lib.lib.fmb_abs_argmax_workspace_bytes(cols))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=x.device)
    cs = x.stride(1) if cols > 1 else rows
    _lib.check(_lib.lib.fmb_abs_argmax(_ptr(x), rows, cols, cs, _t.getFusedType(x.dtype), _ptr(out), _ptr(ws), ws_bytes,
                                       _stream_ptr(x.device)))
    return out


class OMP(Algorithm):

    PARAMETERS = {'numMaxSteps': 0, 'cbStep': None}

    def __init__(self, fmatA, **kwargs):
        if not isinstance(fmatA, Matrix):
            raise TypeError("fmatA must be a fastmat matrix")
        self.fmatA = fmatA
        super(OMP, self).__init__(**kwargs)

    def _process(self, arrB):
        self.arrB, ndim, is_np = _as_device_2d(arrB, self.fmatA)
        if self.numMaxSteps <= 0:
            raise ValueError("OMP would like to do at least one step for you")
        A = self.fmatA
        K = int(self.numMaxSteps)
        N, M, L = A.numRows, A.numCols, self.arrB.shape[1]
        self.numN, self.numM, self.numL = N, M, L
        dev = self.arrB.device
        self.fmatC = A.colNormalized
        ft = _t.promoteTypes(_t.promoteTypes(self.fmatC.fusedType, _t.getFusedType(self.arrB.dtype)), _t.TYPE_FLOAT64)
        tt = _t.getTorchType(ft)
        self.returnType = _t.getNumpyType(ft)
        cplx = tt.is_complex

        # one problem per row of these (L, ...) arrays: every per-step product is a batched matrix product over L
        bT = self.arrB.t().to(tt).contiguous()                      # (L, N)
        resT = bT.clone()                                           # residual, (L, N); its transpose view is column-major (N, L)
        Q = torch.empty((L, K, N), dtype=tt, device=dev)            # orthonormal basis of the support atoms, one per row
        R = torch.zeros((L, K, K), dtype=tt, device=dev)            # atoms = Q^T-combination: a_k = sum_j R[j, k] q_j
        z = torch.zeros((L, K), dtype=tt, device=dev)               # Q^H b
        self.arrSupport = torch.empty((K, L), dtype=torch.long, device=dev)
        onehot = torch.zeros((L, M), dtype=_t.getTorchType(A.fusedType), device=dev).t()     # column-major (M, L) selector
        coef = torch.empty((L, K), dtype=tt, device=dev)
        stream = _stream_ptr(dev)
        cols = torch.arange(L, device=dev)
        prev = None

        def conj(t):
            return t.conj().resolve_conj() if cplx else t

        for self.numStep in range(K):
            k = self.numStep
            # --- selection: strongest correlation of the normalised atoms with the residual
            self.newIndex = abs_argmax(self.fmatC.backward(resT.t()))
            self.arrSupport[k] = self.newIndex
            # --- the picked atoms: one forward apply of a one-hot selector (only L entries of it change per step)
            if prev is not None:
                onehot[prev, cols] = 0
            onehot[self.newIndex, cols] = 1
            prev = self.newIndex
            atom = A.forward(onehot).t().to(tt).contiguous()        # (L, N)
            # --- orthogonalise against the basis so far (twice: classical Gram-Schmidt is not stable enough on its own);
            #     two memory-bound kernels per pass: coef = Q^H v, v -= Q coef
            v = atom
            if k > 0:
                for _ in range(2):
                    _lib.check(_lib.lib.fmb_gs_project(_ptr(Q), K * N, N, k, _ptr(v), N, N, L, _ptr(coef), K, ft, stream))
                    _lib.check(_lib.lib.fmb_gs_subtract(_ptr(Q), K * N, N, k, _ptr(v), N, N, L, _ptr(coef), K, ft, stream))
                    R[:, :k, k] += coef[:, :k]
            rho = torch.linalg.vector_norm(v, dim=1)
            q = v / rho.unsqueeze(1)
            Q[:, k, :] = q
            R[:, k, k] = rho.to(tt)
            # --- coordinate of b along the new direction, residual update with that direction only
            zk = (conj(q) * bT).sum(dim=1)
            z[:, k] = zk
            resT = resT - zk.unsqueeze(1) * q
            self.arrResidual = resT.t()
            self._notify(self.cbStep)
            self._notify(self.cbTrace)
        # coefficients on the support: R x = Q^H b (upper triangular, batched)
        xs = torch.linalg.solve_triangular(R, z.unsqueeze(2), upper=True).squeeze(2)        # (L, K)
        self.arrX = torch.zeros((L, M), dtype=tt, device=dev).t()                           # column-major (M, L)
        self.arrX[self.arrSupport, cols] = xs.t()
        res = self.arrX.reshape(-1) if ndim == 1 else self.arrX
        return _finish(res, is_np)
