"""Small invocations of every 2^20-point operator (for compute-sanitizer memcheck / racecheck runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastmat_b200 as fm
N = 1 << 20
rng = np.random.default_rng(0)
def cr(n): return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
x = torch.view_as_complex(torch.randn((2, N, 2), device='cuda')).t()
ops = [('circulant', fm.Circulant(cr(N)), x), ('fourier', fm.Fourier(N), x),
       ('kron', fm.Kron(fm.Fourier(1024), fm.Fourier(1024)), x),
       ('toeplitz', fm.Toeplitz(cr(N // 2), cr(N // 2 - 1)), x[:N // 2].t().contiguous().t()),
       ('hadamard', fm.Hadamard(20), torch.randn((2, N), device='cuda').t())]
for name, M, xx in ops:
    y = M.forward(xx); z = M.backward(xx)
    torch.cuda.synchronize()
    print(name, 'ok', float(y.abs().sum()) > 0, float(z.abs().sum()) > 0)
