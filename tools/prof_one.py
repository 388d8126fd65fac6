"""Run one operator a few times (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastmat_b200 as fm
what = sys.argv[1] if len(sys.argv) > 1 else 'fourier'
cols = int(sys.argv[2]) if len(sys.argv) > 2 else 16
N = 1 << 20
x = torch.view_as_complex(torch.randn((cols, N, 2), device='cuda')).t()
if what == 'fourier':
    M = fm.Fourier(N)
elif what == 'circulant':
    rng = np.random.default_rng(0)
    M = fm.Circulant((rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64))
elif what == 'hadamard':
    M = fm.Hadamard(20); x = torch.randn((cols, N), device='cuda').t()
for _ in range(3):
    y = M.forward(x)
torch.cuda.synchronize()
print('done', what, cols)
