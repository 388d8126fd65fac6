"""Sparse-recovery solvers driving the apply path (fastmat/algorithms/__init__.py): ISTA, FISTA, OMP, STELA on device tensors."""
from .Algorithm import Algorithm          # noqa: F401
from .ISTA import ISTA, FISTA, ista_step  # noqa: F401
from .OMP import OMP                      # noqa: F401
from .STELA import STELA                  # noqa: F401
