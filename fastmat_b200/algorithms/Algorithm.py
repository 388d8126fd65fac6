"""Solver base class of fastmat_b200.algorithms.

Public behaviour follows fastmat/algorithms/Algorithm.pyx:27-175 (``process(arrB, **params)``, ``updateParameters``,
``snapshot`` / ``trace``, the ``cbResult`` / ``cbTrace`` callbacks, AttributeError for unknown parameters); the mechanics are
this package's own:

* every solver class DECLARES its tunables in ``PARAMETERS`` (name -> default).  The declarations of a class and of its
  bases are merged once per class; ``__init__`` installs the defaults and ``updateParameters`` accepts declared names and
  attributes that already exist - anything else is a typo and raises.
* ``snapshot()`` stores a frozen record of the solver's state (a ``types.SimpleNamespace`` with the instance attributes at
  that moment, the trace itself excluded), so a trace is a plain list of records and holds no solver objects.
"""
from types import SimpleNamespace


class Algorithm(object):

    PARAMETERS = {'cbResult': None, 'cbTrace': None}

    def __init__(self, **params):
        if type(self) is Algorithm:
            raise NotImplementedError("Algorithm baseclass cannot be instantiated.")
        self._trace = []
        for name, default in self.declared_parameters().items():
            setattr(self, name, default)
        self.updateParameters(**params)

    @classmethod
    def declared_parameters(cls):
        """PARAMETERS of the class merged over its bases (most derived declaration wins)."""
        merged = {}
        for klass in reversed(cls.__mro__):
            merged.update(getattr(klass, 'PARAMETERS', {}))
        return merged

    # ---- parameters
    def updateParameters(self, **params):
        declared = self.declared_parameters()
        for name in params:
            if name not in declared and name not in self.__dict__:
                raise AttributeError("Attribute '%s' not defined in %s" % (name, type(self).__name__))
        for name, value in params.items():
            setattr(self, name, value)

    # ---- trace
    @property
    def trace(self):
        return self._trace

    @trace.setter
    def trace(self, records):
        if not isinstance(records, list):
            raise TypeError("Algorithm trace must be a list")
        self._trace = records

    def snapshot(self):
        """Append a record of the current state to the trace (use as ``cbTrace=Algorithm.snapshot``)."""
        state = {k: v for k, v in self.__dict__.items() if k != '_trace'}
        self._trace.append(SimpleNamespace(**state))

    # ---- running
    def process(self, arrB, **params):
        """Run the solver on the right-hand side(s) ``arrB`` (1-D, or 2-D with one problem per column)."""
        self.updateParameters(**params)
        result = self._process(arrB)
        self._notify(self.cbResult)
        return result

    def _process(self, arrB):
        raise NotImplementedError("Algorithm is not implemented yet.")

    def _notify(self, callback):
        return callback(self) if callback is not None else None

    handleCallback = _notify            # name used by the reference (Algorithm.pyx:149-172)


def _as_device_2d(arrB, matrix):
    """Shared input handling of the solvers: numpy or torch in, 2-D CUDA tensor out (+ how to hand the result back)."""
    import numpy as np
    import torch
    is_np = isinstance(arrB, np.ndarray)
    b = torch.from_numpy(np.ascontiguousarray(arrB)) if is_np else arrB
    if not isinstance(b, torch.Tensor):
        raise TypeError("arrB must be a numpy array or a torch tensor")
    if b.ndim > 2 or b.ndim < 1:
        raise ValueError("Only n x m arrays are supported")
    if not b.is_cuda:
        b = b.to(matrix._default_device())
    ndim = b.ndim
    if ndim == 1:
        b = b.reshape(-1, 1)
    return b, ndim, is_np


def _finish(x, is_np):
    return x.cpu().numpy() if is_np else x
