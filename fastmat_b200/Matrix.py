"""Matrix base class: the transform dispatch of fastmat on CUDA tensors.

Mirrors fastmat/Matrix.pyx: ``forward`` (:1844-1910), ``backward`` (:1937-2007) with the input preparation of
``_prepareInputArray`` (:1737-1817) -- dimension check, 1-D <-> 2-D, the dtype plan
fInput -> fInternal = promote(fInput, minType) -> fOutput = promote(fInternal, self.dtype) -- and the operator
interface ``M * x``, ``M * N``, ``M + N``, ``.H / .T / .conj`` (:1209-1269, :1665-1734, :2224-2482).

Arrays are ``torch.Tensor`` on a CUDA device (1-D ``(n,)`` or 2-D ``(n, M)``, any strides; the batch is the second
axis as in the reference).  Results are allocated in the layout class of the input: fastmat's column-major
("fortranStyle", fastmat/core/cmath.pyx:343-386) unless the input is row-major (torch default), in which case the
output is row-major too, so that either way the kernels see one contiguous direction.  A ``numpy.ndarray`` is
accepted as a convenience: it is copied to the current device, transformed there and copied back.

There is no CPU compute path and no dense-matmul bypass (the reference's calibration-driven bypass is inert
without calibration data, fastmat/Matrix.pyx:1406-1409).
"""
import os

import numpy as np
import torch

from . import _lib
from ._lib import lib, check, FORWARD, BACKWARD
from .core import types as _t


# ------------------------------------------------------------------------------------------- helpers
def _stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream


def is_row_major(x):
    """Layout class of a 2-D tensor: True if the batch (second axis) is the contiguous direction."""
    return x.shape[1] > 1 and x.stride(1) == 1 and x.stride(0) != 1


def alloc_out(rows, cols, torch_dtype, device, row_major):
    if row_major:
        return torch.empty((rows, cols), dtype=torch_dtype, device=device)
    return torch.empty((cols, rows), dtype=torch_dtype, device=device).t()      # column-major (rows, cols)


_HOST_CTX = {}
_HOST_STREAMS = {}


def _host_context(dev, rows, step, dtype, row_major):
    """Copy streams (one pair per device) and device staging buffers (per slab shape) of ``apply_host``."""
    key = (dev.index, rows, step, dtype, row_major)
    ctx = _HOST_CTX.get(key)
    if ctx is None:
        if len(_HOST_CTX) >= 8:                      # staging buffers are up to 2 x 64 MiB: keep a handful of shapes
            _HOST_CTX.pop(next(iter(_HOST_CTX)))
        streams = _HOST_STREAMS.get(dev.index)
        if streams is None:
            streams = _HOST_STREAMS[dev.index] = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
        ctx = {'s_in': streams[0], 's_out': streams[1],
               'bufs': [alloc_out(rows, step, dtype, dev, row_major) for _ in range(2)], 'free': [None, None]}
        _HOST_CTX[key] = ctx
    return ctx


def _ptr(t):
    return t.data_ptr() if t.numel() > 0 else None


def plan_apply(plan, direction, x, rows_out, ft_out):
    """One call through the C-ABI: y = A x / A^H x with caller-owned output and workspace (fmb_plan_apply)."""
    M = x.shape[1]
    ft_in = _t.getFusedType(x.dtype)
    y = alloc_out(rows_out, M, _t.getTorchType(ft_out), x.device, is_row_major(x))
    if M == 0 or rows_out == 0:
        return y
    ws_bytes = lib.fmb_plan_workspace_bytes(plan.handle, direction, M, ft_in, ft_out)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes > 0 else None
    check(lib.fmb_plan_apply(plan.handle, direction, x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(), y.stride(0),
                             y.stride(1), M, ft_in, ft_out, ws.data_ptr() if ws is not None else None, ws_bytes,
                             _stream_ptr(x.device)))
    return y


def cast(x, ft_out):
    """Widening cast on the device (own kernel, fmb_cast): fastmat/Matrix.pyx:1814-1815 (_arrForceType)."""
    ft_in = _t.getFusedType(x.dtype)
    if ft_in == ft_out:
        return x
    y = alloc_out(x.shape[0], x.shape[1], _t.getTorchType(ft_out), x.device, is_row_major(x))
    if x.numel():
        check(lib.fmb_cast(x.data_ptr(), x.stride(0), x.stride(1), ft_in, y.data_ptr(), y.stride(0), y.stride(1), ft_out,
                           x.shape[0], x.shape[1], _stream_ptr(x.device)))
    return y


def conjugate(x):
    """fastmat/core/cmath.pyx:744-789 (_conjugate): a conjugated copy; real arrays are returned as they are."""
    if not x.is_complex():
        return x
    y = alloc_out(x.shape[0], x.shape[1], x.dtype, x.device, is_row_major(x))
    if x.numel():
        check(lib.fmb_conjugate(x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(), y.stride(0), y.stride(1), x.shape[0],
                                x.shape[1], _t.getFusedType(x.dtype), _stream_ptr(x.device)))
    return y


def fft_out_type(ft_in, ft_mat):
    """Output type of the FFT-backed operators: promote(input, matrix type), complex; single precision is kept for
    float32 / complex64 / int8 / int16 inputs on a single-precision matrix (documented deviation from the reference,
    which up-casts to complex128 in Product.pyx:215 / Kron.pyx:129; see DESIGN.md "dtype policy")."""
    t = _t.promoteTypes(_t.promoteTypes(ft_in, ft_mat), _t.TYPE_COMPLEX64)
    return t


_REAL_CAST = os.environ.get('FMB_REAL_CAST', '1') != '0'


def fft_in_prepare(x, ft_out, plan=None):
    """Bring x to a dtype the FFT engine reads directly at the output precision (real or complex of that precision).

    The engine reads real input only in its generic run-time-radix kernels; the specialised kernels (power-of-two inner
    length from 64) take complex operands.  For those shapes a real operand is widened first (own kernel, fmb_cast: one
    extra sweep of 12 bytes per element) - 5 to 8 times faster overall than the generic path (FMB_REAL_CAST=0: off)."""
    ft_in = _t.getFusedType(x.dtype)
    if _REAL_CAST and plan is not None and not x.is_complex():
        n = int(plan.info.inner_size)
        if n >= 64 and (n & (n - 1)) == 0:
            return cast(x, ft_out)
    if ft_out == _t.TYPE_COMPLEX64 and ft_in in (_t.TYPE_FLOAT32, _t.TYPE_COMPLEX64):
        return x
    if ft_out == _t.TYPE_COMPLEX128 and ft_in in (_t.TYPE_FLOAT64, _t.TYPE_COMPLEX128):
        return x
    return cast(x, ft_out)


class _Flags(object):
    """fastmat.flags (fastmat/Matrix.pyx:43-48); accepted for compatibility, the bypass does not exist here."""
    bypassAllow = False
    bypassAutoArray = False


flags = _Flags()


# ------------------------------------------------------------------------------------------- Matrix
def _dense_matmul(a, x):
    """a @ x in the numpy / fastmat promotion of the two dtypes (core.types.PROMOTE = np.promote_types, NOT
    torch.promote_types: int32 / int64 with float32 is float64 in numpy and in the reference, fastmat/core/types.pyx:
    443-453).  cuBLAS has no integer GEMM: integer products run on the device as int64 multiply-accumulate (exact, wraps
    like numpy) through a broadcast sum in column chunks."""
    t = _t.getTorchType(_t.promoteTypes(a.dtype, x.dtype))
    if t.is_floating_point or t.is_complex:
        return torch.matmul(a.to(t).resolve_conj(), x.to(t))
    a64, x64 = a.to(torch.int64), x.to(torch.int64)
    out = torch.empty((a.shape[0], x.shape[1]), dtype=torch.int64, device=x.device)
    step = max(1, (1 << 24) // max(1, a.shape[0] * a.shape[1]))
    for c0 in range(0, x.shape[1], step):
        out[:, c0:c0 + step] = (a64.unsqueeze(2) * x64[:, c0:c0 + step].unsqueeze(0)).sum(dim=1)
    return out.to(t)


class Matrix(object):
    """Dense matrix wrapper and base class of every operator (fastmat/Matrix.pyx:1469-1570)."""

    def __init__(self, arrMatrix, **options):
        if isinstance(arrMatrix, np.ndarray):
            arrMatrix = torch.from_numpy(np.ascontiguousarray(arrMatrix))
        if not isinstance(arrMatrix, torch.Tensor):
            raise TypeError("Matrix: Use Sparse() for scipy.sparse matrices; a tensor or ndarray is required.")
        if arrMatrix.ndim != 2:
            raise NotImplementedError("Matrix data array must be 2D.")
        self._array = arrMatrix.detach().clone().to(self._default_device())
        self._initProperties(self._array.shape[0], self._array.shape[1], self._array.dtype, **options)

    # ---- construction helpers
    @staticmethod
    def _default_device():
        if not torch.cuda.is_available():
            raise RuntimeError("fastmat_b200 needs a CUDA device (there is no CPU fallback)")
        return torch.device('cuda', torch.cuda.current_device())

    def _initProperties(self, numRows, numCols, dataType, **options):
        """fastmat/Matrix.pyx:1572-1610."""
        self._numRows = int(numRows)
        self._numCols = int(numCols)
        self._fusedType = _t.getFusedType(dataType)
        self._device_index = torch.cuda.current_device() if torch.cuda.is_available() else None
        self._forceContiguousInput = options.get('forceContiguousInput', False)
        self._widenInputDatatype = options.get('widenInputDatatype', False)
        self._fortranStyle = options.get('fortranStyle', True)
        self._minFusedType = _t.getFusedType(options.get('minType', np.int8))
        self.bypassAllow = False            # accepted and ignored: no dense bypass on the device path
        self.bypassAutoArray = False
        if not hasattr(self, '_content'):
            self._content = ()
        self._tag = options.get('tag', '')
        self._cache = {}

    # ---- basic properties (fastmat/Matrix.pyx:288-360)
    numRows = property(lambda self: self._numRows)
    numCols = property(lambda self: self._numCols)
    shape = property(lambda self: (self._numRows, self._numCols))
    dtype = property(lambda self: _t.getNumpyType(self._fusedType))
    fusedType = property(lambda self: self._fusedType)
    content = property(lambda self: self._content)
    tag = property(lambda self: self._tag)

    def __len__(self):
        return len(self._content)

    def __iter__(self):
        return iter(self._content)

    def __repr__(self):
        return "<%s[%dx%d]:0x%12x>" % (self.__class__.__name__, self.numRows, self.numCols, id(self))

    def __copy__(self):
        return self                         # matrices are immutable (fastmat/Matrix.pyx:341-358)

    def __deepcopy__(self, memo):
        return self

    # ---- overridable transforms
    def _forward(self, x):
        """Dense fallback of the base class itself: array . x (fastmat/Matrix.pyx:1831-1842) via cuBLAS."""
        return _dense_matmul(self._array, x)

    def _backward(self, x):
        a = self._array
        return _dense_matmul(a.conj().t() if a.is_complex() else a.t(), x)

    # ---- input preparation (fastmat/Matrix.pyx:1737-1817)
    def _prepare(self, x, required):
        back_to_numpy = False
        if not isinstance(x, torch.Tensor):
            raise TypeError("Input data must be a torch.Tensor (or numpy.ndarray), got %s" % (type(x).__name__, ))
        if not x.is_cuda:
            raise RuntimeError("Input tensor must live on a CUDA device (fastmat_b200 has no CPU path)")
        if x.ndim < 1 or x.ndim > 2:
            raise ValueError("Input data array must be 1D or 2D")
        if x.shape[0] != required:
            raise ValueError("Mismatch of vector size %d to relevant matrix axis %d" % (x.shape[0], required))
        home = getattr(self, '_device_index', None)
        if home is not None and x.device.index != home:
            # plans (device constants, per-device function attributes) belong to the device that was current at construction
            raise RuntimeError("Input tensor lives on cuda:%d but this matrix was created on cuda:%d" % (x.device.index, home))
        ndim = x.ndim
        if x.is_complex() and x.is_conj():
            x = x.resolve_conj()
        if x.is_neg():
            x = x.resolve_neg()
        if ndim == 1:
            x = x.reshape(required, 1)
        f_in = _t.getFusedType(x.dtype)
        f_internal = _t.promoteTypes(f_in, self._minFusedType)
        f_out = _t.promoteTypes(f_internal, self._fusedType)
        if self._widenInputDatatype:
            f_internal = f_out
        if f_internal != f_in:
            x = cast(x, f_internal)
        return x, ndim, back_to_numpy

    @staticmethod
    def _finish(y, ndim, back_to_numpy):
        if ndim == 1:
            y = y.reshape(-1)
        if back_to_numpy:
            return y.cpu().numpy()
        return y

    # ---- public transforms
    def forward(self, arrX):
        """y = A x.  fastmat/Matrix.pyx:1844-1910.  CUDA tensor in -> CUDA tensor out; a numpy array (host data) is
        streamed through the device by ``apply_host`` and comes back as a numpy array."""
        if isinstance(arrX, np.ndarray):
            return self.apply_host(arrX, backward=False)
        x, ndim, np_out = self._prepare(arrX, self.numCols)
        return self._finish(self._forward(x), ndim, np_out)

    def backward(self, arrX):
        """y = A^H x.  fastmat/Matrix.pyx:1937-2007."""
        if isinstance(arrX, np.ndarray):
            return self.apply_host(arrX, backward=True)
        x, ndim, np_out = self._prepare(arrX, self.numRows)
        return self._finish(self._backward(x), ndim, np_out)

    def apply_host(self, arrX, backward=False, out=None, chunk_bytes=64 << 20, stats=None):
        """Host buffers in, host buffers out: the call a user with numpy data makes (``M.forward(ndarray)``).

        The column batch is cut into slabs of ``chunk_bytes`` (64 MiB: the fill and drain of the pipeline cost one slab each;
        measured on B200 / PCIe 5: 5.46 k columns/s against 5.07 k with 256 MiB slabs, tools/e2e_chunks.py); slab k+1 is
        copied host->device while slab k is transformed and slab k-1 is copied device->host.  The two copy streams and the
        two device staging buffers are created once per (device, slab shape) and reused by every later call on any matrix
        (round 1 created streams per call).  ``arrX`` may be a numpy array or a CPU torch tensor (pinned memory makes the
        copies asynchronous); ``out`` an optional preallocated CPU tensor / array of the result shape.  Returns the same kind
        of object that came in.  ``stats`` (a dict) receives the bytes moved in each direction.
        """
        is_np = isinstance(arrX, np.ndarray)
        xh = torch.from_numpy(arrX) if is_np else arrX
        if not isinstance(xh, torch.Tensor) or xh.is_cuda:
            raise TypeError("apply_host expects a numpy array or a CPU tensor")
        _t.getFusedType(xh.dtype)
        required = self.numRows if backward else self.numCols
        if xh.ndim < 1 or xh.ndim > 2:
            raise ValueError("Input data array must be 1D or 2D")
        if xh.shape[0] != required:
            raise ValueError("Mismatch of vector size %d to relevant matrix axis %d" % (xh.shape[0], required))
        ndim = xh.ndim
        x2 = xh.reshape(required, 1) if ndim == 1 else xh
        M = x2.shape[1]
        dev = self._default_device()
        rows_out = self.numCols if backward else self.numRows
        step = max(1, min(max(M, 1), chunk_bytes // max(1, required * x2.element_size())))
        cur = torch.cuda.current_stream(dev)
        ctx = _host_context(dev, required, step, x2.dtype, is_row_major(x2))
        s_in, s_out, bufs, free_ev = ctx['s_in'], ctx['s_out'], ctx['bufs'], ctx['free']
        # a previous call may still be draining on the copy streams: order this call behind it
        s_in.wait_stream(cur)
        out_h = None
        if out is not None:
            out_h = torch.from_numpy(out) if isinstance(out, np.ndarray) else out
            out_h = out_h.reshape(rows_out, -1)
        h2d = d2h = 0
        for k, c0 in enumerate(range(0, max(M, 1), step)):
            c1 = min(M, c0 + step)
            if c1 <= c0:
                break
            b = k % 2
            with torch.cuda.stream(s_in):
                if free_ev[b] is not None:
                    s_in.wait_event(free_ev[b])
                xd = bufs[b][:, :c1 - c0]
                xd.copy_(x2[:, c0:c1], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(s_in)
            h2d += xd.numel() * xd.element_size()
            cur.wait_event(ready)
            xp, _, _ = self._prepare(xd, required)
            yd = self._backward(xp) if backward else self._forward(xp)
            free_ev[b] = torch.cuda.Event()
            free_ev[b].record(cur)
            if out_h is None:
                out_h = torch.empty((M, rows_out), dtype=yd.dtype, pin_memory=xh.is_pinned()).t() if not is_row_major(x2) \
                    else torch.empty((rows_out, M), dtype=yd.dtype, pin_memory=xh.is_pinned())
            with torch.cuda.stream(s_out):
                s_out.wait_event(free_ev[b])
                out_h[:, c0:c1].copy_(yd, non_blocking=True)
                yd.record_stream(s_out)
            d2h += yd.numel() * yd.element_size()
        s_out.synchronize()
        if stats is not None:
            stats['h2d_bytes'] = h2d
            stats['d2h_bytes'] = d2h
            stats['chunks'] = (max(M, 1) + step - 1) // step
        if out_h is None:
            out_h = torch.empty((rows_out, M), dtype=_t.getTorchType(self._fusedType))
        res = out_h.reshape(-1) if ndim == 1 else out_h
        return res.numpy() if is_np else res

    # ---- views (fastmat/Matrix.pyx:1209-1269)
    @property
    def H(self):
        if 'H' not in self._cache:
            self._cache['H'] = self._getH()
        return self._cache['H']

    @property
    def T(self):
        if 'T' not in self._cache:
            self._cache['T'] = self._getT()
        return self._cache['T']

    @property
    def conj(self):
        if 'conj' not in self._cache:
            self._cache['conj'] = self._getConj()
        return self._cache['conj']

    def _getH(self):
        return Hermitian(self)

    def _getT(self):
        return Transpose(self)

    def _getConj(self):
        return getConjugate(self)

    # ---- operator interface (fastmat/Matrix.pyx:1665-1734)
    def __add__(self, other):
        from .Sum import Sum
        if isinstance(other, Matrix):
            return Sum(self, other)
        raise TypeError("Not an addition of fastmat matrices.")

    __radd__ = __add__

    def __sub__(self, other):
        from .Sum import Sum
        from .Product import Product
        if isinstance(other, Matrix):
            return Sum(self, Product(other, -1))
        raise TypeError("Not a subtraction of fastmat matrices.")

    def __rsub__(self, other):
        from .Sum import Sum
        from .Product import Product
        if isinstance(other, Matrix):
            return Sum(other, Product(self, -1))
        raise TypeError("Not a subtraction of fastmat matrices.")

    def __mul__(self, other):
        from .Product import Product
        if isinstance(other, (torch.Tensor, np.ndarray)) and getattr(other, 'ndim', 0) >= 1:
            return self.forward(other)
        if isinstance(other, Matrix) or np.isscalar(other) or (isinstance(other, (torch.Tensor, np.ndarray)) and other.ndim == 0):
            return Product(self, other)
        raise TypeError("Invalid product term for fastmat Matrix.")

    def __rmul__(self, other):
        from .Product import Product
        if np.isscalar(other) or isinstance(other, Matrix):
            return Product(other, self)
        raise TypeError("Invalid product term for fastmat Matrix.")

    def __truediv__(self, other):
        from .Product import Product
        if np.isscalar(other):
            return Product(self, 1.0 / other)
        raise TypeError("Only scalars allowed as divisors")

    __div__ = __truediv__

    # ---- dense views, built from the transform itself (fastmat/Matrix.pyx:361-657)
    def getArray(self):
        """fastmat/Matrix.pyx:376-387: the dense matrix as forward(eye)."""
        if getattr(self, '_array', None) is None:
            eye = torch.eye(self.numCols, dtype=_t.getTorchType(self._fusedType), device=self._default_device())
            self._array = self.forward(eye)
        return self._array

    array = property(lambda self: self.getArray())

    def getCols(self, idx):
        if np.isscalar(idx):
            return self.getCol(int(idx))
        idx = (idx if isinstance(idx, torch.Tensor) else torch.as_tensor(np.asarray(idx))).to(self._default_device()).long().reshape(-1)
        sel = torch.zeros((self.numCols, idx.numel()), dtype=_t.getTorchType(self._fusedType), device=idx.device)
        sel[idx, torch.arange(idx.numel(), device=idx.device)] = 1
        return self.forward(sel)

    def getCol(self, idx):
        """fastmat/Matrix.pyx:454-458: forward(e_idx)."""
        if idx < 0 or idx >= self.numCols:
            raise ValueError("Column index exceeds matrix dimensions.")
        e = torch.zeros(self.numCols, dtype=_t.getTorchType(self._fusedType), device=self._default_device())
        e[idx] = 1
        return self.forward(e)

    def getRows(self, idx):
        if np.isscalar(idx):
            return self.getRow(int(idx))
        idx = (idx if isinstance(idx, torch.Tensor) else torch.as_tensor(np.asarray(idx))).to(self._default_device()).long().reshape(-1)
        sel = torch.zeros((self.numRows, idx.numel()), dtype=_t.getTorchType(self._fusedType), device=idx.device)
        sel[idx, torch.arange(idx.numel(), device=idx.device)] = 1
        return self.backward(sel).conj().resolve_conj().t()

    def getRow(self, idx):
        if idx < 0 or idx >= self.numRows:
            raise ValueError("Row index exceeds matrix dimensions.")
        e = torch.zeros(self.numRows, dtype=_t.getTorchType(self._fusedType), device=self._default_device())
        e[idx] = 1
        r = self.backward(e)
        return r.conj().resolve_conj() if r.is_complex() else r

    def __getitem__(self, tplIdx):
        if not isinstance(tplIdx, tuple) or len(tplIdx) != 2:
            raise ValueError("Matrix element access requires two indices.")
        i, j = tplIdx
        if np.isscalar(i) and np.isscalar(j):
            return self.getCol(int(j))[int(i)]
        return self.getArray()[i, j]

    # ---- norms / gram (fastmat/Matrix.pyx:1008-1206), generic versions through the transform
    @property
    def gram(self):
        from .Product import Product
        if 'gram' not in self._cache:
            self._cache['gram'] = self._getGram()
        return self._cache['gram']

    def _getGram(self):
        from .Product import Product
        return Product(self.H, self)

    @property
    def colNorms(self):
        if 'colNorms' not in self._cache:
            self._cache['colNorms'] = self._getColNorms()
        return self._cache['colNorms']

    def _getColNorms(self):
        """fastmat/Matrix.pyx:1048-1088: forward on chunks of unit vectors, 2-norm of the result columns."""
        dev = self._default_device()
        out = torch.empty(self.numCols, dtype=torch.float64, device=dev)
        chunk = max(1, min(self.numCols, (1 << 24) // max(1, self.numRows)))
        # in double: the norms feed step sizes / normalisations of the solvers (the reference's FFT classes are complex128)
        tt = _t.getTorchType(_t.promoteTypes(self._fusedType, _t.TYPE_FLOAT64))
        for c0 in range(0, self.numCols, chunk):
            c1 = min(self.numCols, c0 + chunk)
            sel = torch.zeros((self.numCols, c1 - c0), dtype=tt, device=dev)
            sel[torch.arange(c0, c1, device=dev), torch.arange(c1 - c0, device=dev)] = 1
            out[c0:c1] = torch.linalg.vector_norm(self.forward(sel), dim=0).to(torch.float64)
        return out

    @property
    def rowNorms(self):
        if 'rowNorms' not in self._cache:
            self._cache['rowNorms'] = self._getRowNorms()
        return self._cache['rowNorms']

    def _getRowNorms(self):
        return self.H.colNorms

    @property
    def colNormalized(self):
        from .Product import Product
        from .Diag import Diag
        if 'colNormalized' not in self._cache:
            n = self.colNorms
            if bool((n == 0).any()):
                raise ValueError("Normalization: Matrix has zero-norm column.")
            self._cache['colNormalized'] = Product(self, Diag(1.0 / n))
        return self._cache['colNormalized']

    @property
    def rowNormalized(self):
        from .Product import Product
        from .Diag import Diag
        if 'rowNormalized' not in self._cache:
            n = self.rowNorms
            if bool((n == 0).any()):
                raise ValueError("Normalization: Matrix has zero-norm row.")
            self._cache['rowNormalized'] = Product(Diag(1.0 / n), self)
        return self._cache['rowNormalized']

    @property
    def largestSingularValue(self):
        if 'lsv' not in self._cache:
            self._cache['lsv'] = self._getLargestSingularValue()
        return self._cache['lsv']

    def _getLargestSingularValue(self, maxSteps=100, relEps=1e-13):
        """sqrt of the largest eigenvalue of A^H A by Lanczos iteration on the device (full re-orthogonalisation, Ritz
        value of the small tridiagonal matrix on the host).  The reference calls scipy's ARPACK ``svds(k=1)``
        (fastmat/Matrix.pyx:895-919) -- the same Krylov method -- on ``scipyLinearOperator``; every step here is one
        forward and one backward apply through the C-ABI."""
        dev = self._default_device()
        tt = _t.getTorchType(_t.promoteTypes(self._fusedType, _t.TYPE_FLOAT64))
        g = torch.Generator(device=dev)
        g.manual_seed(1234)
        q = torch.randn(self.numCols, dtype=torch.float64, device=dev, generator=g).to(tt)
        q = q / torch.linalg.vector_norm(q)
        maxSteps = max(1, min(maxSteps, self.numCols))
        Q = torch.zeros((maxSteps, self.numCols), dtype=tt, device=dev)
        alphas, betas = [], []
        theta = 0.0
        for j in range(maxSteps):
            Q[j] = q
            w = self.backward(self.forward(q)).to(tt)
            alphas.append(float(torch.real(torch.vdot(q, w))))
            # full re-orthogonalisation against the whole Krylov basis (twice is enough)
            for _ in range(2):
                w = w - (Q[:j + 1].conj() @ w) @ Q[:j + 1]
            beta = float(torch.linalg.vector_norm(w))
            T = np.diag(alphas) + np.diag(betas, 1) + np.diag(betas, -1)
            new = float(np.linalg.eigvalsh(T)[-1])
            done = j >= 2 and abs(new - theta) <= relEps * abs(new)
            theta = new
            if done or beta <= 1e-14 * max(abs(theta), 1e-300) or j + 1 == maxSteps:
                break
            betas.append(beta)
            q = w / beta
        return float(np.sqrt(max(theta, 0.0)))

    @property
    def largestEigenValue(self):
        """Largest-magnitude eigenvalue of a square matrix by power iteration on the device, Rayleigh quotient as the
        estimate (what fastmat/Matrix.pyx:678-760 computes with numpy; every step is one forward apply through the C-ABI).
        Cached like the reference's property."""
        if 'lev' not in self._cache:
            self._cache['lev'] = self._getLargestEigenValue()
        return self._cache['lev']

    def _getLargestEigenValue(self, maxSteps=10000, relEps=1e-12):
        if self.numRows != self.numCols:
            raise ValueError("largestEigenValue: Matrix must be square.")
        dev = self._default_device()
        tt = _t.getTorchType(_t.promoteTypes(self._fusedType, _t.TYPE_FLOAT64))
        g = torch.Generator(device=dev)
        g.manual_seed(4321)
        v = torch.randn(self.numCols, dtype=torch.float64, device=dev, generator=g).to(tt)
        v = v / torch.linalg.vector_norm(v)
        lam = 0.0
        for step in range(maxSteps):
            w = self.forward(v).to(tt)
            new = torch.vdot(v, w)                                       # Rayleigh quotient (||v|| = 1)
            nrm = float(torch.linalg.vector_norm(w))
            if nrm == 0.0:
                return 0.0
            v = w / nrm
            if step % 8 == 7:                                            # one host read-back every eight steps
                cur = complex(new.item()) if tt.is_complex else float(new.item())
                if abs(cur - lam) <= relEps * abs(cur):
                    lam = cur
                    break
                lam = cur
        else:
            lam = complex(new.item()) if tt.is_complex else float(new.item())
        return lam.real if isinstance(lam, complex) and abs(lam.imag) <= 1e-12 * abs(lam) else lam

    @property
    def scipyLinearOperator(self):
        """scipy.sparse.linalg.LinearOperator over this matrix for host (numpy) vectors, as fastmat/Matrix.pyx:977-1006:
        matvec / matmat = forward, rmatvec / rmatmat = backward; host arrays are streamed through the device by
        apply_host."""
        from scipy.sparse.linalg import LinearOperator
        return LinearOperator(shape=self.shape, dtype=self.dtype,
                              matvec=lambda x: self.apply_host(np.asarray(x).reshape(-1)),
                              rmatvec=lambda x: self.apply_host(np.asarray(x).reshape(-1), backward=True),
                              matmat=lambda x: self.apply_host(np.asarray(x)),
                              rmatmat=lambda x: self.apply_host(np.asarray(x), backward=True))

    def reference(self):
        """Dense reference of the matrix, built without the fast transform where a class overrides _reference."""
        return self._reference()

    def _reference(self):
        return self.getArray()


# ------------------------------------------------------------------------------------------- views
class Hermitian(Matrix):
    """fastmat/Matrix.pyx:2224-2300: swaps forward and backward."""

    def __init__(self, matrix):
        if not isinstance(matrix, Matrix):
            raise TypeError("Hermitian: Not a fastmat Matrix")
        self._content = (matrix, )
        self._initProperties(matrix.numCols, matrix.numRows, matrix.fusedType)

    def _getH(self):
        return self._content[0]

    def _getT(self):
        return self._content[0].conj

    def _getConj(self):
        return self._content[0].T

    def _forward(self, x):
        return self._content[0].backward(x)

    def _backward(self, x):
        return self._content[0].forward(x)

    def _reference(self):
        r = self._content[0].reference()
        return (r.conj() if r.is_complex() else r).t()


class Conjugate(Matrix):
    """fastmat/Matrix.pyx:2326-2412: conj(x) -> nested -> conj."""

    def __init__(self, matrix):
        if not isinstance(matrix, Matrix):
            raise TypeError("Conjugate: Not a fastmat Matrix")
        self._content = (matrix, )
        self._initProperties(matrix.numRows, matrix.numCols, matrix.fusedType)

    def _getConj(self):
        return self._content[0]

    def _getH(self):
        return self._content[0].T

    def _getT(self):
        return self._content[0].H

    def _forward(self, x):
        y = self._content[0].forward(conjugate(x))
        return conjugate(y)

    def _backward(self, x):
        y = self._content[0].backward(conjugate(x))
        return conjugate(y)

    def _reference(self):
        r = self._content[0].reference()
        return r.conj().resolve_conj() if r.is_complex() else r


def getConjugate(matrix):
    """fastmat/Matrix.pyx:2304-2322: real matrices are their own conjugate."""
    return Conjugate(matrix) if _t.isComplex(matrix.fusedType) else matrix


class Transpose(Hermitian):
    """fastmat/Matrix.pyx:2417-2482: T = H(conj(M))."""

    def __init__(self, matrix):
        if not isinstance(matrix, Matrix):
            raise TypeError("Transpose: Not a fastmat Matrix")
        self._inner = matrix
        super(Transpose, self).__init__(getConjugate(matrix))

    def _getT(self):
        return self._inner

    def _getH(self):
        return self._inner.conj

    def _getConj(self):
        return self._inner.H
