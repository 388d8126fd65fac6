"""A/B of the index (gather / scatter) kernels on the LFSRCirculant order-20 row: FMB_EW_TILED=1 (column-tiled permute
kernel) against FMB_EW_TILED=0 (flat grid-stride kernels).  Run once per setting: the switch is read once per process."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastmat_b200 as fm                        # noqa: E402

N = (1 << 20) - 1
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = fm.LFSRCirculant((1 << 20) | (1 << 3) | 1, 1)
g = torch.Generator(device='cuda').manual_seed(1)
x = torch.randn((cols, N), dtype=torch.float32, device='cuda', generator=g).t()


def timed(fn, k=5, w=2):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


H = L.content[0]
d = torch.zeros((cols, 1 << 20), dtype=torch.float32, device='cuda').t()
from fastmat_b200.Matrix import plan_apply      # noqa: E402
from fastmat_b200._lib import FORWARD, BACKWARD  # noqa: E402
ft = 4
print('FMB_EW_TILED=%s cols=%d' % (os.environ.get('FMB_EW_TILED', '1'), cols))
print('  forward   %.3f ms' % timed(lambda: L.forward(x)))
print('  backward  %.3f ms' % timed(lambda: L.backward(x)))
print('  scatter   %.3f ms (zero + scatter)' % timed(lambda: plan_apply(L._scatterFlip, BACKWARD, x, 1 << 20, ft)))
print('  hadamard  %.3f ms' % timed(lambda: H.forward(d)))
print('  gather    %.3f ms' % timed(lambda: plan_apply(L._gather, FORWARD, d, N, ft)))
y = L.forward(x[:, :4].contiguous())
print('  checksum %.6e' % float(y.double().sum()))
