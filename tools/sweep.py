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
C = fm.Circulant(c); F = fm.Fourier(N)
tc = timed(lambda: C.forward(x)); tf = timed(lambda: F.forward(x))
gb = 16.0 * N * cols / 1e9
print("env SLAB_MB=%s TILE=%s WIDE=%s | circ %.2f ms %.0f GB/s (%.1f%%) | fourier %.2f ms %.0f GB/s (%.1f%%)" % (
    os.environ.get('FMB_SLAB_MB'), os.environ.get('FMB_TILE_ELEMS'), os.environ.get('FMB_WIDE_T'),
    tc, gb / tc * 1e3, gb / tc * 1e3 / 65.539, tf, gb / tf * 1e3, gb / tf * 1e3 / 65.539))
