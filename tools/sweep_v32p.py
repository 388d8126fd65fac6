"""Tuning sweep (experiments only): Fourier / Circulant 2^20 under the FMB_V32P_* knobs of the current environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastmat_b200 as fm
what, cols = sys.argv[1], int(sys.argv[2])
N = 1 << 20
x = torch.view_as_complex(torch.randn((cols, N, 2), device='cuda')).t()
def timed(f, k=5):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
if what == 'fourier':
    M = fm.Fourier(N)
elif what == 'kron':
    M = fm.Kron(fm.Fourier(1024), fm.Fourier(1024))
elif what == 'toeplitz':
    rng = np.random.default_rng(0)
    nt = N // 2
    M = fm.Toeplitz((rng.standard_normal(nt) + 1j * rng.standard_normal(nt)).astype(np.complex64),
                    (rng.standard_normal(nt - 1) + 1j * rng.standard_normal(nt - 1)).astype(np.complex64))
    x = x[:nt].t().contiguous().t()
else:
    rng = np.random.default_rng(0)
    M = fm.Circulant((rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64))
ts = [timed(lambda: M.forward(x)) for _ in range(3)]
gb = 16.0 * N * cols / 1e9
env = ' '.join('%s=%s' % (k[4:], v) for k, v in sorted(os.environ.items()) if k.startswith('FMB_'))
print('%-9s %4d cols | %-40s | %.3f ms (min of %s) %.1f%%' % (what, cols, env, min(ts), ' '.join('%.3f' % t for t in ts), gb / min(ts) * 1e3 / 6550.1 * 100), flush=True)
