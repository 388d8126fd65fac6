"""Where does OMP on the config-5 operator spend its time?  torch.profiler table of one solve (experiments)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastmat_b200 as fm
from torch.profiler import profile, ProfilerActivity
n, m, k, L = 1 << 18, 1 << 16, 32, int(sys.argv[1]) if len(sys.argv) > 1 else 128
rng = np.random.default_rng(2026)
rows = np.sort(rng.choice(n, m, replace=False))
d = np.exp(2j * np.pi * rng.random(n)).astype(np.complex64)
A = fm.Product(fm.Partial(fm.Fourier(n), rows=rows), fm.Diag(d))
x = np.zeros((L, n), dtype=np.complex64)
for c in range(L):
    idx = rng.choice(n, k, replace=False)
    x[c, idx] = (2 + rng.random(k)) * np.exp(2j * np.pi * rng.random(k))
xd = torch.from_numpy(x).cuda().t()
b = A.forward(xd)
omp = fm.algorithms.OMP(A, numMaxSteps=k)
omp.process(b, numMaxSteps=2)
omp.numMaxSteps = k
torch.cuda.synchronize(); t0 = time.perf_counter()
res = omp.process(b)
torch.cuda.synchronize(); print("OMP k=%d L=%d: %.3f s, support exact %s" % (k, L, time.perf_counter() - t0, bool(torch.equal(res != 0, xd != 0))))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    omp.process(b, numMaxSteps=8)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
