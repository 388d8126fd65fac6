import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import fastmat_b200 as fm
n = int(sys.argv[1]); cols = (1 << 25) // n
x = torch.view_as_complex(torch.randn((cols, n, 2), dtype=torch.float32, device='cuda')).t()
c = (np.random.default_rng(0).standard_normal(n) + 0j).astype(np.complex64)
C = fm.Circulant(c); F = fm.Fourier(n)
for _ in range(3):
    y = C.forward(x); z = F.forward(x)
torch.cuda.synchronize()
