"""Partial: row / column sub-selection of a matrix.  Mirrors fastmat/Partial.pyx.

forward (fastmat/Partial.pyx:268-280): scatter x into the selected columns of a zero buffer, nested forward,
gather the selected rows; backward (:282-294) mirrored.  Gather / scatter are exact index kernels
(``fmb_partial_plan_create``).  Contiguous leading ranges on a Fourier / Circulant / Toeplitz are absorbed by those
classes' own load / store masks instead (see Toeplitz / Circulant).
"""
import ctypes
import warnings

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, FORWARD, BACKWARD
from .Matrix import Matrix, plan_apply
from .Circulant import _to_host
from .core import types as _t


def _check_selection(sel, size, name):
    """fastmat/Partial.pyx:141-166."""
    if sel is None:
        return None
    sel = np.array(_to_host(sel))
    if sel.dtype == bool:
        sel = np.arange(size)[sel]
    elif np.issubdtype(sel.dtype, np.integer):
        pass
    else:
        raise TypeError("Partial: Type of %s indices must be int or bool." % (name, ))
    if sel.ndim != 1:
        sel = sel.reshape(-1)
    bounded = ((len(sel) != size) or (np.sum(sel - np.arange(size)) != 0)) and np.any((sel >= size) | (sel < 0))
    if bounded:
        raise ValueError("Partial: A %s index exceeds matrix dimensions." % (name, ))
    return sel.astype(np.int64)


def _make_plan(sel, total):
    if sel is None:
        return None
    idx = np.ascontiguousarray(sel, dtype=np.int64)
    h = ctypes.c_void_p()
    check(lib.fmb_partial_plan_create(ctypes.byref(h), idx.ctypes.data_as(ctypes.c_void_p) if idx.size else None,
                                      int(idx.size), int(total)))
    return _lib.Plan(h)


class Partial(Matrix):

    def __init__(self, mat, **options):
        if not isinstance(mat, Matrix):
            raise TypeError("Partial: fastmat Matrix required.")
        self._content = (mat, )
        if 'N' in options:
            warnings.warn('N=~ is deprecated in Partial.__init__(). Use rows=~.', FutureWarning)
            options['rows'] = options['N']
        if 'M' in options:
            warnings.warn('M=~ is deprecated in Partial.__init__() Use cols=~.', FutureWarning)
            options['cols'] = options['M']
        self._rowSelection = _check_selection(options.get('rows', None), mat.numRows, 'row')
        self._colSelection = _check_selection(options.get('cols', None), mat.numCols, 'col')
        self._rowPlan = _make_plan(self._rowSelection, mat.numRows)
        self._colPlan = _make_plan(self._colSelection, mat.numCols)
        self._initProperties(len(self._rowSelection) if self._rowSelection is not None else mat.numRows,
                             len(self._colSelection) if self._colSelection is not None else mat.numCols,
                             mat.fusedType, **options)

    rowSelection = property(lambda self: self._rowSelection)
    colSelection = property(lambda self: self._colSelection)

    def __repr__(self):
        if type(self) is Partial:
            m = self._content[0]
            return "<%s[%dx%d](%s[%dx%d]):0x%12x>" % (self.__class__.__name__, self.numRows, self.numCols,
                                                      m.__class__.__name__, m.numRows, m.numCols, id(self))
        return super(Partial, self).__repr__()

    def _forward(self, x):
        ft = _t.getFusedType(x.dtype)
        if self._colPlan is not None:
            x = plan_apply(self._colPlan, BACKWARD, x, self._content[0].numCols, ft)      # zero + scatter
        y = self._content[0].forward(x)
        if self._rowPlan is not None:
            y = plan_apply(self._rowPlan, FORWARD, y, self._numRows, _t.getFusedType(y.dtype))   # gather
        return y

    def _backward(self, x):
        ft = _t.getFusedType(x.dtype)
        if self._rowPlan is not None:
            x = plan_apply(self._rowPlan, BACKWARD, x, self._content[0].numRows, ft)
        y = self._content[0].backward(x)
        if self._colPlan is not None:
            y = plan_apply(self._colPlan, FORWARD, y, self._numCols, _t.getFusedType(y.dtype))
        return y

    # norms of a pure row / column selection follow from the nested matrix (fastmat/Partial.pyx:234-250)
    def _getColNorms(self):
        if self._rowSelection is not None:
            return super(Partial, self)._getColNorms()
        n = self._content[0].colNorms
        return n if self._colSelection is None else n[torch.from_numpy(self._colSelection).to(n.device)]

    def _getRowNorms(self):
        if self._colSelection is not None:
            return super(Partial, self)._getRowNorms()
        n = self._content[0].rowNorms
        return n if self._rowSelection is None else n[torch.from_numpy(self._rowSelection).to(n.device)]

    def _reference(self):
        full = self._content[0].reference()
        if self._rowSelection is not None:
            full = full[torch.from_numpy(self._rowSelection).to(full.device)]
        if self._colSelection is not None:
            full = full[:, torch.from_numpy(self._colSelection).to(full.device)]
        return full
