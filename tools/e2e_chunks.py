"""Experiment: host-buffer throughput of Circulant(2^20).apply_host against the chunk size (PCIe pipeline fill / drain)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastmat_b200 as fm
N, cols = 1 << 20, 256
rng = np.random.default_rng(0)
C = fm.Circulant((rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64))
xh = torch.empty((cols, N), dtype=torch.complex64, pin_memory=True).t()
xh.copy_(torch.view_as_complex(torch.randn((cols, N, 2))).t())
yh = torch.empty((cols, N), dtype=torch.complex64, pin_memory=True).t()
ref = None
for mb in (256, 128, 64, 32, 16, 256):
    C.apply_host(xh, out=yh, chunk_bytes=mb << 20)
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); C.apply_host(xh, out=yh, chunk_bytes=mb << 20); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    chk = float(yh.real.double().sum())
    if ref is None: ref = chk
    print('chunk %4d MiB: %.1f ms  %.0f columns/s  (checksum %s)' % (mb, min(ts) * 1e3, cols / min(ts), 'same' if chk == ref else 'DIFFERENT'), flush=True)
