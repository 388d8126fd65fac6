"""Block matrix.  Mirrors fastmat/Blocks.pyx (_forwardC :210-248: row-slices of x, accumulate row[c].forward(x_c))."""
import torch

from .Matrix import Matrix, cast, alloc_out, is_row_major
from .core import types as _t


def _add(a, b):
    """a + b in the numpy / fastmat promotion of the two dtypes (not torch's: int32 + float32 is float64 there)."""
    t = _t.getTorchType(_t.promoteTypes(a.dtype, b.dtype))
    return a.to(t) + b.to(t)


class Blocks(Matrix):

    def __init__(self, arrMatrices, **options):
        if not isinstance(arrMatrices, (list, tuple)):
            raise ValueError("Blocks: Not a nested list of fastmat matrices.")
        if len(arrMatrices) < 1:
            raise ValueError("Blocks: Contains no matrices.")
        self._rows = []
        self._rowSize = []
        self._colSize = []
        ft = _t.TYPE_INT8
        numRows = 0
        for rr, row in enumerate(arrMatrices):
            if not isinstance(row, (list, tuple)):
                raise ValueError("Blocks: Not a nested list of fastmat matrices.")
            if rr == 0:
                self._numN = len(row)
                self._colSize = [m.numCols for m in row if isinstance(m, Matrix)]
            if len(row) != self._numN:
                raise ValueError("Blocks.row(%d) has incompatible number of entries" % (rr, ))
            for cc, term in enumerate(row):
                if not isinstance(term, Matrix):
                    raise TypeError("Blocks: Not a fastmat Matrix at (%d, %d)." % (rr, cc))
                if cc == 0:
                    height = term.numRows
                elif term.numRows != height:
                    raise ValueError("Blocks.row(%d): Heights of blocks differ." % (rr, ))
                if term.numCols != self._colSize[cc]:
                    raise ValueError("Blocks.col(%d): Widths of blocks differ." % (cc, ))
                ft = _t.promoteTypes(ft, term.fusedType)
            self._rows.append(tuple(row))
            self._rowSize.append(height)
            numRows += height
        self._content = tuple(m for row in self._rows for m in row)
        self._initProperties(numRows, sum(self._colSize), ft, **options)

    def _forward(self, x):
        outs = []
        for row in self._rows:
            acc = None
            c0 = 0
            for term in row:
                y = term.forward(x[c0:c0 + term.numCols, :])
                acc = y if acc is None else _add(acc, y)
                c0 += term.numCols
            outs.append(acc)
        return self._stack(outs, x)

    def _backward(self, x):
        outs = [None] * self._numN
        r0 = 0
        for rr, row in enumerate(self._rows):
            xs = x[r0:r0 + self._rowSize[rr], :]
            for cc, term in enumerate(row):
                y = term.backward(xs)
                outs[cc] = y if outs[cc] is None else _add(outs[cc], y)
            r0 += self._rowSize[rr]
        return self._stack(outs, x)

    def _stack(self, outs, x):
        ft = _t.promoteTypes(x.dtype, self._fusedType)
        for y in outs:
            ft = _t.promoteTypes(ft, y.dtype)
        total = sum(y.shape[0] for y in outs)
        res = alloc_out(total, x.shape[1], _t.getTorchType(ft), x.device, is_row_major(x))
        r0 = 0
        for y in outs:
            res[r0:r0 + y.shape[0], :] = y
            r0 += y.shape[0]
        return res

    def _reference(self):
        rows = []
        for row in self._rows:
            refs = [m.reference() for m in row]
            t = refs[0].dtype
            for r in refs:
                t = _t.promoteTorch(t, r.dtype)
            rows.append(torch.cat([r.to(t) for r in refs], dim=1))
        t = rows[0].dtype
        for r in rows:
            t = _t.promoteTorch(t, r.dtype)
        return torch.cat([r.to(t) for r in rows], dim=0)
