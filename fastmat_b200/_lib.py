"""ctypes binding of libfastmat_b200.so (the C-ABI declared in include/fastmat_b200.h).

There is no CPU fallback: importing this module without the compiled CUDA library raises ImportError, and every
entry point that needs a device fails with RuntimeError when none is present.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('FMB_LIB_PATH') or os.path.join(_HERE, 'lib', 'libfastmat_b200.so')   # override: A/B experiments

FMB_OK, FMB_ERR_VALUE, FMB_ERR_TYPE, FMB_ERR_CUDA, FMB_ERR_NOTIMPL, FMB_ERR_WORKSPACE = 0, -1, -2, -3, -4, -5
FORWARD, BACKWARD = 0, 1

c_i64 = ctypes.c_int64
c_vp = ctypes.c_void_p


class PlanInfo(ctypes.Structure):
    _fields_ = [('kind', ctypes.c_int32), ('num_rows', c_i64), ('num_cols', c_i64), ('inner_size', c_i64),
                ('bluestein', c_i64), ('passes_fwd', ctypes.c_int32), ('slab_cols', ctypes.c_int32)]


# every symbol include/fastmat_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    'fmb_last_error': (ctypes.c_char_p, []),
    'fmb_version': (ctypes.c_int, []),
    'fmb_device_info': (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)] * 3 + [ctypes.POINTER(ctypes.c_size_t)]),
    'fmb_find_optimal_fft_size': (c_i64, [c_i64, ctypes.c_int]),
    'fmb_fft_complexity': (ctypes.c_float, [c_i64]),
    'fmb_lfsr_order': (ctypes.c_int, [ctypes.c_uint32]),
    'fmb_lfsr_period': (c_i64, [ctypes.c_uint32, ctypes.c_uint32]),
    'fmb_lfsr_sequences': (ctypes.c_int, [ctypes.c_uint32, ctypes.c_uint32, c_i64, c_vp, c_vp, c_vp]),
    'fmb_fourier_plan_create': (ctypes.c_int, [ctypes.POINTER(c_vp), c_i64, ctypes.c_int, ctypes.c_int]),
    'fmb_circulant_plan_create': (ctypes.c_int, [ctypes.POINTER(c_vp), c_vp, c_i64, ctypes.c_int, ctypes.c_int]),
    'fmb_toeplitz_plan_create': (ctypes.c_int, [ctypes.POINTER(c_vp), c_vp, c_i64, c_vp, c_i64, ctypes.c_int, ctypes.c_int]),
    'fmb_hadamard_plan_create': (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int]),
    'fmb_diag_plan_create': (ctypes.c_int, [ctypes.POINTER(c_vp), c_vp, ctypes.c_int, c_i64]),
    'fmb_partial_plan_create': (ctypes.c_int, [ctypes.POINTER(c_vp), c_vp, c_i64, c_i64]),
    'fmb_kron_fourier_plan_create': (ctypes.c_int, [ctypes.POINTER(c_vp), c_vp, ctypes.c_int]),
    'fmb_plan_info_get': (ctypes.c_int, [c_vp, ctypes.POINTER(PlanInfo)]),
    'fmb_plan_workspace_bytes': (c_i64, [c_vp, ctypes.c_int, c_i64, ctypes.c_int, ctypes.c_int]),
    'fmb_plan_apply': (ctypes.c_int, [c_vp, ctypes.c_int, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, ctypes.c_int,
                                      ctypes.c_int, c_vp, c_i64, c_vp]),
    'fmb_plan_destroy': (ctypes.c_int, [c_vp]),
    'fmb_conjugate': (ctypes.c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_i64, ctypes.c_int, c_vp]),
    'fmb_ista_step': (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_double, ctypes.c_double, ctypes.c_int, c_vp]),
    'fmb_cast': (ctypes.c_int, [c_vp, c_i64, c_i64, ctypes.c_int, c_vp, c_i64, c_i64, ctypes.c_int, c_i64, c_i64, c_vp]),
    'fmb_abs_argmax_workspace_bytes': (c_i64, [c_i64]),
    'fmb_abs_argmax': (ctypes.c_int, [c_vp, c_i64, c_i64, c_i64, ctypes.c_int, c_vp, c_vp, c_i64, c_vp]),
    'fmb_gs_project': (ctypes.c_int, [c_vp, c_i64, c_i64, ctypes.c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, ctypes.c_int, c_vp]),
    'fmb_gs_subtract': (ctypes.c_int, [c_vp, c_i64, c_i64, ctypes.c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, ctypes.c_int, c_vp]),
    'fmb_launch_count': (c_i64, []),
}


def bind(cdll):
    """Attach restype/argtypes for every declared symbol; raises AttributeError if one is not exported."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    return cdll


def load(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            "fastmat_b200: compiled CUDA library not found at %s -- build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or ./build.sh (there is no CPU fallback)" % path)
    return bind(ctypes.CDLL(path))


lib = load()


def check(rc):
    """Map a C status to the exception classes the reference raises (fastmat/Matrix.pyx:1772-1782, core/types.pyx:159)."""
    if rc == FMB_OK:
        return
    msg = lib.fmb_last_error().decode('utf-8', 'replace')
    if rc in (FMB_ERR_VALUE, FMB_ERR_WORKSPACE):
        raise ValueError(msg)
    if rc == FMB_ERR_TYPE:
        raise TypeError(msg)
    if rc == FMB_ERR_NOTIMPL:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


class Plan(object):
    """Owner of one fmb_plan (destroyed with the object)."""

    def __init__(self, handle):
        self.handle = handle
        info = PlanInfo()
        check(lib.fmb_plan_info_get(handle, ctypes.byref(info)))
        self.info = info

    def __del__(self):
        h, self.handle = getattr(self, 'handle', None), None
        if h:
            try:
                lib.fmb_plan_destroy(h)
            except Exception:           # interpreter shutdown
                pass
