"""Zero matrix.  Mirrors fastmat/Zero.pyx."""
import numpy as np
import torch

from .Matrix import Matrix, alloc_out, is_row_major


class Zero(Matrix):

    def __init__(self, numRows, numCols, **options):
        self._initProperties(int(numRows), int(numCols), np.int8, **options)

    def _forward(self, x):
        return alloc_out(self._numRows, x.shape[1], x.dtype, x.device, is_row_major(x)).zero_()

    def _backward(self, x):
        return alloc_out(self._numCols, x.shape[1], x.dtype, x.device, is_row_major(x)).zero_()

    def _reference(self):
        return torch.zeros((self._numRows, self._numCols), dtype=torch.int8, device=self._default_device())
