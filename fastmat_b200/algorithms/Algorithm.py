"""Algorithm base class (fastmat/algorithms/Algorithm.pyx:27-175): parameter handling, callbacks, trace."""
from copy import copy


class Algorithm(object):

    def __init__(self):
        if type(self) is Algorithm:
            raise NotImplementedError("Algorithm baseclass cannot be instantiated.")

    # ---- callbacks / trace (Algorithm.pyx:48-75)
    cbTrace = None
    cbResult = None
    _trace = None

    @property
    def trace(self):
        if self._trace is None:
            self._trace = []
        return self._trace

    @trace.setter
    def trace(self, value):
        if not isinstance(value, list):
            raise TypeError("Algorithm trace must be a list")
        self._trace = value

    def updateParameters(self, **kwargs):
        """Algorithm.pyx:86-108: setattr every keyword; unknown attributes raise AttributeError."""
        if getattr(self, '_attributes', None) is None:
            self._attributes = kwargs.copy()
        for key, value in kwargs.items():
            if not hasattr(self, key) and (self._attributes is not None and key not in self._attributes):
                raise AttributeError("Attribute '%s' not defined in %s" % (key, self.__class__.__name__))
            setattr(self, key, value)

    def process(self, arrB, **kwargs):
        """Algorithm.pyx:110-126."""
        self.updateParameters(**kwargs)
        arrResult = self._process(arrB)
        self.handleCallback(self.cbResult)
        return arrResult

    def _process(self, arrB):
        raise NotImplementedError("Algorithm is not implemented yet.")

    def snapshot(self):
        """Algorithm.pyx:136-147: append a copy of the current state (without the trace) to the trace."""
        trace, self._trace = self._trace, []
        if trace is None:
            trace = []
        trace.append(copy(self))
        self._trace = trace

    def handleCallback(self, callback):
        if callback is not None:
            return callback(self)
        return None


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
