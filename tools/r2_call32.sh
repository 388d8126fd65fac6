#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
import fastmat_b200 as fm
def timed(f, k=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
for n in (2**14, 2**15):
    x = torch.view_as_complex(torch.randn((1024, n, 2), device='cuda')).t()
    F = fm.Fourier(n); C = fm.Circulant(np.random.default_rng(0).standard_normal(n).astype(np.complex64))
    tf, tc = timed(lambda: F.forward(x)), timed(lambda: C.forward(x))
    gb = 16.0 * n * 1024 / 1e9
    print("n=%d x 1024: fourier %.3f ms (%.3f of roofline), circulant %.3f ms (%.3f)" % (n, tf, gb / tf * 1e3 / 6449.7, tc, gb / tc * 1e3 / 6449.7))
PY
