"""Tuning sweep (experiments only): time Circulant/Fourier 2^20 for the current env-var knobs."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastmat_b200 as fm
N = 1 << 20
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(0)
c = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
x = torch.view_as_complex(torch.randn((cols, N, 2), device='cuda')).t()
def timed(f, k=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
if len(sys.argv) > 2 and sys.argv[2] == 'hadamard':
    H = fm.Hadamard(20)
    xf = torch.randn((cols * 2, N), device='cuda').t()
    th = timed(lambda: H.forward(xf))
    gbh = 8.0 * N * cols * 2 / 1e9
    print("env FWHT_SLAB_MB=%s | hadamard o20 f32 %d cols: %.2f ms %.0f GB/s (%.1f%%)" % (os.environ.get('FMB_FWHT_SLAB_MB'), cols * 2, th, gbh / th * 1e3, gbh / th * 1e3 / 65.539))
    xi = torch.randint(-2**31, 2**31 - 1, (64, N), device='cuda', dtype=torch.int32).t()
    z = H.forward(H.forward(xi))
    print("involution int32 exact:", bool(torch.equal(z, xi * N)))
    sys.exit(0)
C = fm.Circulant(c); F = fm.Fourier(N)
tc = timed(lambda: C.forward(x)); tf = timed(lambda: F.forward(x))
yc = C.forward(x); chk = float(yc.real.double().sum() + 3 * yc.imag.double().sum()); chk2 = float(yc.abs().double().sum())
gb = 16.0 * N * cols / 1e9
print("env PIPE=%s PIPE_MB=%s chk %.6e %.8e SLAB_MB=%s TILE=%s WIDE=%s | circ %.2f ms %.0f GB/s (%.1f%%) | fourier %.2f ms %.0f GB/s (%.1f%%)" % (
    os.environ.get('FMB_PIPE_STREAMS'), os.environ.get('FMB_PIPE_MB'), chk, chk2, os.environ.get('FMB_SLAB_MB'), os.environ.get('FMB_TILE_ELEMS'), os.environ.get('FMB_WIDE_T'),
    tc, gb / tc * 1e3, gb / tc * 1e3 / 65.539, tf, gb / tf * 1e3, gb / tf * 1e3 / 65.539))
