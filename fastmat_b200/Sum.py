"""Sum of matrices.  Mirrors fastmat/Sum.pyx (_forwardC :151-176: out = c0.forward(x); out += ci.forward(x))."""
import torch

from .Matrix import Matrix, cast
from .core import types as _t


class Sum(Matrix):

    def __init__(self, *matrices, **options):
        terms = []

        def add(items):
            for m in items:
                if not isinstance(m, Matrix):
                    raise TypeError("Sum: Term is not a Matrix.")
                if isinstance(m, Sum):
                    add(m.content)
                else:
                    terms.append(m)
        add(matrices)
        if len(terms) < 1:
            raise ValueError("Sum: No terms given.")
        numRows, numCols = terms[0].numRows, terms[0].numCols
        ft = _t.TYPE_INT8
        for m in terms:
            if m.numRows != numRows or m.numCols != numCols:
                raise ValueError("Sum: Term dimension mismatch: " + repr(m))
            ft = _t.promoteTypes(ft, m.fusedType)
        self._content = tuple(terms)
        self._initProperties(numRows, numCols, ft, **options)

    def _accumulate(self, outs, x):
        ft = _t.promoteTypes(x.dtype, self._fusedType)
        for y in outs:
            ft = _t.promoteTypes(ft, y.dtype)
        acc = cast(outs[0], ft)
        if acc is outs[0]:
            acc = acc.clone()
        for y in outs[1:]:
            acc += y                      # torch promotes the addend; the accumulator keeps the output type
        return acc

    def _forward(self, x):
        return self._accumulate([m.forward(x) for m in self._content], x)

    def _backward(self, x):
        return self._accumulate([m.backward(x) for m in self._content], x)

    def _reference(self):
        arr = None
        for m in self._content:
            r = m.reference()
            arr = r if arr is None else arr.to(_t.promoteTorch(arr.dtype, r.dtype)) + r
        return arr
