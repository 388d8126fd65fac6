"""Real operands on the FFT-backed operators: widened to complex before the specialised kernels (default) against the
engine's real-input generic path (FMB_REAL_CAST=0).  Circulant / Fourier 2^20 x 256 float32 columns, Circulant 4096."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def worker():
    import numpy as np, torch
    sys.path.insert(0, ROOT)
    import fastmat_b200 as fm
    rng = np.random.default_rng(2)
    for n, cols in ((2 ** 20, 256), (4096, 16384), (2 ** 16, 1024)):
        x = torch.randn((cols, n), dtype=torch.float32, device='cuda').t()
        c = rng.standard_normal(n).astype(np.float32)
        for name, op in (('circulant', fm.Circulant(c)), ('fourier', fm.Fourier(n))):
            y = op.forward(x)
            ref = torch.fft.fft(x[:, :4].to(torch.complex128), dim=0)
            if name == 'circulant':
                ref = torch.fft.ifft(ref * torch.fft.fft(torch.from_numpy(c).cuda().to(torch.complex128)).unsqueeze(1), dim=0)
            err = float((y[:, :4] - ref).abs().max() / ref.abs().max())
            for _ in range(2): y = op.forward(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): y = op.forward(x)
            e1.record(); torch.cuda.synchronize()
            print('%-9s n=%7d cols=%5d float32 in -> %s: %.3f ms  err %.1e' % (name, n, cols, str(y.dtype)[6:], e0.elapsed_time(e1) / 5, err), flush=True)
if __name__ == '__main__':
    if len(sys.argv) > 1: worker()
    else:
        for v in ('1', '0'):
            print('== FMB_REAL_CAST=%s' % v, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), 'w'], env=dict(os.environ, FMB_REAL_CAST=v))
