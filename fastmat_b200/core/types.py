"""The eight fastmat data types on torch tensors.

Mirrors fastmat/core/types.pyx: ftype ids 0..7 (types.pxd:45-55), the promotion table (= np.promote_types over
the eight types, types.pyx:443-453, documented in doc/architecture/types.rst:65-89) and safeTypeExpansion
(types.pyx:378-394).  Unlike the reference (types.pyx:76-111), dtypes outside the eight are rejected with
TypeError instead of being approximated by item size.
"""
import numpy as np
import torch

TYPE_INT8, TYPE_INT16, TYPE_INT32, TYPE_INT64, TYPE_FLOAT32, TYPE_FLOAT64, TYPE_COMPLEX64, TYPE_COMPLEX128 = range(8)

_NUMPY = [np.dtype(t) for t in ('int8', 'int16', 'int32', 'int64', 'float32', 'float64', 'complex64', 'complex128')]
_TORCH = [torch.int8, torch.int16, torch.int32, torch.int64, torch.float32, torch.float64, torch.complex64,
          torch.complex128]
_FROM_TORCH = {t: i for i, t in enumerate(_TORCH)}
_FROM_NUMPY = {t: i for i, t in enumerate(_NUMPY)}

# promotion table, computed the way the reference does it at import time
PROMOTE = [[_FROM_NUMPY[np.promote_types(a, b)] for b in _NUMPY] for a in _NUMPY]

__all__ = ['getFusedType', 'getNumpyType', 'getTorchType', 'promoteTypes', 'safeTypeExpansion', 'isComplex', 'isInteger',
           'isFloat', 'getTypeEps', 'TYPE_INT8', 'TYPE_INT16', 'TYPE_INT32', 'TYPE_INT64', 'TYPE_FLOAT32', 'TYPE_FLOAT64',
           'TYPE_COMPLEX64', 'TYPE_COMPLEX128']


def getFusedType(obj):
    """ftype id of a torch dtype / numpy dtype / tensor / ndarray / python type (types.pyx getFusedType)."""
    if isinstance(obj, int) and not isinstance(obj, bool) and 0 <= obj < 8:
        return obj
    if isinstance(obj, torch.Tensor):
        obj = obj.dtype
    if isinstance(obj, torch.dtype):
        try:
            return _FROM_TORCH[obj]
        except KeyError:
            raise TypeError("Not a fastmat fused type: %s" % (obj, ))
    if isinstance(obj, np.ndarray):
        obj = obj.dtype
    if obj is int:
        obj = np.int64
    elif obj is float:
        obj = np.float64
    elif obj is complex:
        obj = np.complex128
    try:
        return _FROM_NUMPY[np.dtype(obj)]
    except (KeyError, TypeError):
        raise TypeError("Not a fastmat fused type: %s" % (obj, ))


def getNumpyType(obj):
    return _NUMPY[getFusedType(obj)].type


def getTorchType(obj):
    return _TORCH[getFusedType(obj)]


def promoteTypes(a, b):
    """ftype id of the promotion of two types (any spelling)."""
    return PROMOTE[getFusedType(a)][getFusedType(b)]


def safeTypeExpansion(dtype):
    """types.pyx:378-394: int8/int16 -> float32, int32/int64 -> float64, others unchanged (returns an ftype id)."""
    t = getFusedType(dtype)
    if t in (TYPE_INT8, TYPE_INT16):
        return TYPE_FLOAT32
    if t in (TYPE_INT32, TYPE_INT64):
        return TYPE_FLOAT64
    return t


def isComplex(obj):
    return getFusedType(obj) >= TYPE_COMPLEX64


def isInteger(obj):
    return getFusedType(obj) <= TYPE_INT64


def isFloat(obj):
    return getFusedType(obj) in (TYPE_FLOAT32, TYPE_FLOAT64)


def getTypeEps(obj):
    """types.pyx:213-233: machine eps of the type (0 for integers)."""
    t = getFusedType(obj)
    if t <= TYPE_INT64:
        return 0.0
    return float(np.finfo(_NUMPY[t]).eps)


def promoteTorch(a, b):
    """torch dtype of the numpy / fastmat promotion of two torch (or numpy) dtypes - use this, never
    torch.promote_types (which differs from numpy: int32 / int64 with float32 gives float32 there)."""
    return getTorchType(promoteTypes(a, b))
