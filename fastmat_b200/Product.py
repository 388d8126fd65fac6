"""Product of matrices and scalars.  Mirrors fastmat/Product.pyx.

Nested products are flattened and scalars folded into one factor (fastmat/Product.pyx:90-114); forward applies the
factors right to left after promoting the input to promote(x, dtype) (:205-221), backward left to right with the
conjugated scalar (:223-240).
"""
import numpy as np
import torch

from .Matrix import Matrix, cast
from .core import types as _t


# Python scalars are *weakly* typed, as in numpy >= 2 (NEP 50): their value never decides the operator's dtype.  A python
# int / float / complex only lifts the KIND of the matrix factors' common type (integer -> int64 / float64 / complex128,
# the reference's np.array(scalar).dtype; real floating -> the complex type of the same precision) and never its
# precision, so M * 0.5, M * 0.1, M / 2 and M / 3 all have the dtype class of M, and a single-precision operator stays
# single.  numpy scalars keep their own dtype (reference: fastmat/Product.pyx:108-114).
_WEAK_INT, _WEAK_FLOAT, _WEAK_COMPLEX = 1, 2, 3


def _scalar_type(f):
    """(fastmat type or None, weak kind or 0) of a scalar factor."""
    if isinstance(f, np.generic):
        return _t.getFusedType(f.dtype), 0
    if isinstance(f, (bool, int)):
        if not (np.iinfo(np.int64).min <= int(f) <= np.iinfo(np.int64).max):
            raise TypeError("Product: integer scalar out of range.")
        return None, _WEAK_INT
    if isinstance(f, float):
        return None, _WEAK_FLOAT
    if isinstance(f, complex):
        return None, _WEAK_COMPLEX
    raise TypeError("Product: Term is neither scalar nor Matrix.")


def _apply_weak(ft, weak):
    """Lift the matrix factors' common type `ft` by the strongest weak python scalar kind seen."""
    if weak == 0:
        return ft
    if _t.isInteger(ft):
        return {_WEAK_INT: _t.TYPE_INT64, _WEAK_FLOAT: _t.TYPE_FLOAT64, _WEAK_COMPLEX: _t.TYPE_COMPLEX128}[weak]
    if weak == _WEAK_COMPLEX and not _t.isComplex(ft):
        return _t.TYPE_COMPLEX64 if ft == _t.TYPE_FLOAT32 else _t.TYPE_COMPLEX128
    return ft


class Product(Matrix):

    def __init__(self, *matrices, **options):
        debug = options.get('debug', False)
        scalar = [1]
        factors = []
        ft = [_t.TYPE_INT8]
        weak = [0]

        def add(items):
            for f in items:
                if isinstance(f, torch.Tensor) and f.ndim == 0:
                    f = f.item()
                if isinstance(f, np.ndarray) and f.ndim == 0:
                    f = f[()]
                if isinstance(f, Matrix):
                    if isinstance(f, Product):
                        if f._scalar != 1:
                            scalar[0] = scalar[0] * f._scalar
                        add(f.content)
                    else:
                        factors.append(f)
                    ft[0] = _t.promoteTypes(ft[0], f.fusedType)
                elif np.isscalar(f):
                    if f != 1:
                        scalar[0] = scalar[0] * f
                    st, wk = _scalar_type(f)
                    if st is not None:
                        ft[0] = _t.promoteTypes(ft[0], st)
                    weak[0] = max(weak[0], wk)
                else:
                    raise TypeError("Product: Term is neither scalar nor Matrix.")

        add(matrices)
        dtype = _apply_weak(ft[0], weak[0])
        expansion = options.get('typeExpansion', _t.safeTypeExpansion(dtype))
        if expansion is not None:
            dtype = _t.promoteTypes(dtype, expansion)
        if len(factors) < 1:
            raise ValueError("Product has no terms.")
        numRows, numCols = factors[0].numRows, factors[0].numCols
        for ii in range(1, len(factors)):
            if factors[ii].numRows != numCols:
                raise ValueError("Product: Dimension mismatch for term %d [%dx%d]" % (ii, numRows, numCols))
            numCols = factors[ii].numCols
        self._scalar = np.asarray(scalar[0]).astype(_t.getNumpyType(dtype))[()]     # :143
        self._content = tuple(factors)
        self._initProperties(numRows, numCols, dtype, **options)
        if debug:
            print("fastmat_b200.Product instance %12x containing:" % (id(self), ))
            for ii, f in enumerate(factors):
                print("  [%d]: %s" % (ii, repr(f)))

    def _scale_in(self, x, scalar):
        ft = _t.promoteTypes(x.dtype, self._fusedType)
        x = cast(x, ft)
        if scalar != 1:
            if _t.isComplex(ft):
                s = complex(scalar)
            elif _t.isInteger(ft):
                s = int(scalar)
            else:
                s = float(scalar)
            x = x * s
        return x

    def _forward(self, x):
        r = self._scale_in(x, self._scalar)
        for f in reversed(self._content):
            r = f.forward(r)
        return r

    def _backward(self, x):
        s = np.conj(self._scalar) if np.iscomplexobj(self._scalar) else self._scalar
        r = self._scale_in(x, s)
        for f in self._content:
            r = f.backward(r)
        return r

    # Exact shortcuts the generic unit-vector sweep (fastmat/Matrix.pyx:1048-1088) does not need: a trailing Diag scales
    # the columns, a leading Diag the rows, a scalar both.  Needed to make OMP's colNormalized affordable for the
    # compressed-sensing operator Product(Partial(Fourier), Diag) with 2^18 columns.
    def _getColNorms(self):
        from .Diag import Diag
        c = self._content
        if len(c) >= 1 and isinstance(c[-1], Diag):
            rest = c[0].colNorms if len(c) == 2 else (Product(*c[:-1]).colNorms if len(c) > 2 else None)
            d = c[-1].colNorms
            n = d if rest is None else rest * d
            return n * abs(self._scalar)
        if len(c) == 1:
            return c[0].colNorms * abs(self._scalar)
        return super(Product, self)._getColNorms()

    def _getRowNorms(self):
        from .Diag import Diag
        c = self._content
        if len(c) >= 1 and isinstance(c[0], Diag):
            rest = c[1].rowNorms if len(c) == 2 else (Product(*c[1:]).rowNorms if len(c) > 2 else None)
            d = c[0].rowNorms
            n = d if rest is None else rest * d
            return n * abs(self._scalar)
        if len(c) == 1:
            return c[0].rowNorms * abs(self._scalar)
        return super(Product, self)._getRowNorms()

    def _reference(self):
        arr = None
        for f in self._content:
            r = f.reference()
            r = r.to(_t.getTorchType(_t.promoteTypes(_t.promoteTypes(r.dtype, self._fusedType), _t.TYPE_FLOAT32)))
            if arr is None:
                arr = r
            else:
                t = _t.promoteTorch(arr.dtype, r.dtype)
                arr = arr.to(t) @ r.to(t)
        if self._scalar != 1:
            arr = arr * (complex(self._scalar) if np.iscomplexobj(self._scalar) else float(self._scalar))
        return arr
