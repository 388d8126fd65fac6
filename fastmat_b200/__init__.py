"""fastmat_b200 -- the structured-matrix apply path of fastmat as hand-written sm_100a CUDA.

Same class names and call semantics as EMS-TU-Ilmenau/fastmat for the hot path (Matrix.forward / backward of
Fourier, Circulant, Toeplitz, Hadamard, Kron and the Partial / Product / Diag / Sum / Blocks / .H / .T / .conj
glue), operating on torch CUDA tensors through the C-ABI in include/fastmat_b200.h.  No CPU fallback.
"""
from ._lib import lib as _clib, LIB_PATH                           # noqa: F401  (fails loudly if the library is missing)
from .Matrix import Matrix, Hermitian, Conjugate, Transpose, flags   # noqa: F401
from .Fourier import Fourier                                       # noqa: F401
from .Circulant import Circulant                                   # noqa: F401
from .Toeplitz import Toeplitz                                     # noqa: F401
from .Hadamard import Hadamard                                     # noqa: F401
from .Diag import Diag                                             # noqa: F401
from .Partial import Partial                                       # noqa: F401
from .Product import Product                                       # noqa: F401
from .Sum import Sum                                               # noqa: F401
from .LFSRCirculant import LFSRCirculant                         # noqa: F401
from .Kron import Kron                                             # noqa: F401
from .Blocks import Blocks                                         # noqa: F401
from .BlockDiag import BlockDiag                                   # noqa: F401
from .Permutation import Permutation                               # noqa: F401
from .Eye import Eye                                               # noqa: F401
from .Zero import Zero                                             # noqa: F401
from . import core                                                 # noqa: F401
from . import algorithms                                           # noqa: F401

__version__ = '0.1.0'


def launch_count():
    """Kernel launches issued by the library in this process (bench.py's gpu_launches)."""
    return int(_clib.fmb_launch_count())
