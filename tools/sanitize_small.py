"""Small invocations of every kernel family (for compute-sanitizer memcheck / racecheck runs):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py
2^20-point operators in the single-column and in the pipelined-slab schedule (TMA-fed first pass, pruned Toeplitz passes),
the persistent kernel, the row-major route (transposes), pass length 128, FWHT, the solver kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastmat_b200 as fm
N = 1 << 20
rng = np.random.default_rng(0)
def cr(n): return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 13
x = torch.view_as_complex(torch.randn((cols, N, 2), device='cuda')).t()
ops = [('circulant', fm.Circulant(cr(N)), x), ('fourier', fm.Fourier(N), x),
       ('kron', fm.Kron(fm.Fourier(1024), fm.Fourier(1024)), x),
       ('toeplitz', fm.Toeplitz(cr(N // 2), cr(N // 2 - 1)), x[:N // 2].t().contiguous().t()),
       ('hadamard', fm.Hadamard(20), torch.randn((cols, N), device='cuda').t()),
       ('circulant row-major', fm.Circulant(cr(N)), x[:, :3].contiguous()),
       ('circulant 2^14', fm.Circulant(cr(1 << 14)), x[:1 << 14].t().contiguous().t()),
       ('fourier 2^15 c128', fm.Fourier(1 << 15), x[:1 << 15].t().contiguous().t().to(torch.complex128)),
       ('single column', fm.Circulant(cr(N)), x[:, :1].t().contiguous().t())]
# single-kernel routes (L <= 4096: 16-value kernel, column- and row-major, ragged tails; 1024: 32-value kernel incl. pruned and
# masked Toeplitz), 8192 = 64 x 128, int32 FWHT (64-value first pass, 128-bit strided pass), real operands
def cm(n, m): return torch.view_as_complex(torch.randn((m, n, 2), device='cuda')).t()
for n in (64, 128, 1024, 4096, 8192):
    xs_ = cm(n, 70 + n % 7)
    ops.append(('fourier %d' % n, fm.Fourier(n), xs_))
    ops.append(('circulant %d' % n, fm.Circulant(cr(n)), xs_))
    ops.append(('circulant %d row-major' % n, fm.Circulant(cr(n)), xs_.contiguous()))
    ops.append(('toeplitz %d' % (n // 2), fm.Toeplitz(cr(n // 2), cr(n // 2 - 1)), xs_[:n // 2]))
ops.append(('toeplitz 500 x 512', fm.Toeplitz(cr(500), cr(511)), None))
ops.append(('hadamard int32', fm.Hadamard(20), torch.randint(-2 ** 31, 2 ** 31 - 1, (cols, N), dtype=torch.int32, device='cuda').t()))
ops.append(('hadamard 24', fm.Hadamard(24), torch.randn((2, 1 << 24), device='cuda').t()))
ops.append(('circulant real operand', fm.Circulant(rng.standard_normal(N).astype(np.float32)), torch.randn((5, N), device='cuda').t()))
for name, M, xx in ops:
    if xx is None:                                       # non-square: forward and backward take different operands
        y = M.forward(cm(512, 40)); z = M.backward(cm(500, 40))
        torch.cuda.synchronize()
        print(name, 'ok', float(y.abs().sum()) > 0, float(z.abs().sum()) > 0)
        continue
    y = M.forward(xx); z = M.backward(xx)
    torch.cuda.synchronize()
    print(name, 'ok', float(y.abs().sum()) > 0, float(z.abs().sum()) > 0)
n, m, k, L = 1 << 12, 1 << 10, 8, 5
rows = np.sort(rng.choice(n, m, replace=False))
A = fm.Product(fm.Partial(fm.Fourier(n), rows=rows), fm.Diag(np.exp(2j * np.pi * rng.random(n)).astype(np.complex64)))
xs = np.zeros((n, L), dtype=np.complex64)
for c in range(L):
    xs[rng.choice(n, k, replace=False), c] = 3.0
b = A.forward(torch.from_numpy(np.ascontiguousarray(xs.T)).cuda().t())
r = fm.algorithms.OMP(A, numMaxSteps=k).process(b)
torch.cuda.synchronize()
print('omp ok', bool(torch.equal(r != 0, torch.from_numpy(np.ascontiguousarray(xs.T)).cuda().t() != 0)))
r = fm.algorithms.ISTA(A, numLambda=1.0, numMaxSteps=5).process(b)
torch.cuda.synchronize()
print('ista ok')
