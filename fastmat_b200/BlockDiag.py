"""Block diagonal matrix.  Mirrors fastmat/BlockDiag.pyx (_forwardC :145-161)."""
import torch

from .Matrix import Matrix, alloc_out, is_row_major
from .core import types as _t


class BlockDiag(Matrix):

    def __init__(self, *matrices, **options):
        if len(matrices) < 1:
            raise ValueError("BlockDiag: No matrices given.")
        ft = _t.TYPE_INT8
        numRows = numCols = 0
        for m in matrices:
            if not isinstance(m, Matrix):
                raise ValueError("BlockDiag: Term is not a fastmat Matrix.")
            numRows += m.numRows
            numCols += m.numCols
            ft = _t.promoteTypes(ft, m.fusedType)
        self._content = tuple(matrices)
        self._initProperties(numRows, numCols, ft, **options)

    def _apply(self, x, backward):
        outs = []
        i0 = 0
        for m in self._content:
            n_in = m.numRows if backward else m.numCols
            xs = x[i0:i0 + n_in, :]
            outs.append(m.backward(xs) if backward else m.forward(xs))
            i0 += n_in
        ft = _t.promoteTypes(x.dtype, self._fusedType)
        for y in outs:
            ft = _t.promoteTypes(ft, y.dtype)
        res = alloc_out(sum(y.shape[0] for y in outs), x.shape[1], _t.getTorchType(ft), x.device, is_row_major(x))
        r0 = 0
        for y in outs:
            res[r0:r0 + y.shape[0], :] = y
            r0 += y.shape[0]
        return res

    def _forward(self, x):
        return self._apply(x, False)

    def _backward(self, x):
        return self._apply(x, True)

    def _reference(self):
        refs = [m.reference() for m in self._content]
        t = refs[0].dtype
        for r in refs:
            t = _t.promoteTorch(t, r.dtype)
        return torch.block_diag(*[r.to(t) for r in refs])
