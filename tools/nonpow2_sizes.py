"""Circulant / Toeplitz of sizes that are not powers of two: the padded length the reference's planner picks (run-time-radix
kernels) against the next power of two (specialised kernels).  FMB_POW2_PAD=0 / 1, one subprocess each."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def worker():
    import numpy as np, torch
    sys.path.insert(0, ROOT)
    import fastmat_b200 as fm
    from oracle import fastmat_oracle as orc
    rng = np.random.default_rng(2)
    for n in (100, 1000, 3000, 6144, 41000, 100000, 786432):
        cols = max(8, (1 << 25) // n)
        x = torch.view_as_complex(torch.randn((cols, n, 2), dtype=torch.float32, device='cuda')).t()
        c = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        for name, op, xin in (('circulant', fm.Circulant(c), x), ('toeplitz', fm.Toeplitz(c, c[1:][::-1].copy()), x)):
            y = op.forward(xin)
            xs = xin[:, :2].cpu().numpy()
            ref = orc.circulant_forward(c, xs) if name == 'circulant' else orc.toeplitz_forward(c, c[1:][::-1], xs)
            err = float(np.abs(y[:, :2].cpu().numpy() - ref).max() / (np.linalg.norm(c) * np.linalg.norm(xs, axis=0).max() * np.log2(n)))
            for _ in range(2): y = op.forward(xin)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): y = op.forward(xin)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print('%-9s n=%7d cols=%6d inner=%8d  %8.3f ms  %.3f of peak  err %.1e' % (name, n, cols, int(op._plan.info.inner_size), ms, 16.0 * n * cols / ms / 1e6 / 6449.7, err), flush=True)
if __name__ == '__main__':
    if len(sys.argv) > 1: worker()
    else:
        for v in ('0', '1'):
            print('== FMB_POW2_PAD=%s' % v, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), 'w'], env=dict(os.environ, FMB_POW2_PAD=v))
