"""Time the single-kernel sizes (L <= 8192) with CUDA events: Fourier / Circulant / Toeplitz forward, complex64, a column
batch of 2^26 elements (512 MiB in, 512 MiB out: larger than L2), specialised route against the generic kernel
(FMB_NO_FAST1=1).  Prints ms and the fraction of the measured HBM peak.  ROWMAJOR=1: batch-contiguous operands.

    python tools/small_sizes.py            # both routes, one subprocess each
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker():
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import fastmat_b200 as fm
    peak = 6449.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rng = np.random.default_rng(1)
    for n in (128, 256, 512, 1024, 2048, 4096, 8192):
        cols = (1 << 26) // n
        if os.environ.get("ROWMAJOR", "0") == "1":
            x = torch.view_as_complex(torch.randn((n, cols, 2), dtype=torch.float32, device="cuda"))
        else:
            x = torch.view_as_complex(torch.randn((cols, n, 2), dtype=torch.float32, device="cuda")).t()
        c = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        nt = n // 2
        ops = [("fourier", fm.Fourier(n), x), ("circulant", fm.Circulant(c), x),
               ("toeplitz", fm.Toeplitz(c[:nt], c[nt:2 * nt - 1]), x[:nt])]
        for name, op, xin in ops:
            for _ in range(3):
                y = op.forward(xin)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                y = op.forward(xin)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            nbytes = 8.0 * (xin.shape[0] * cols + y.shape[0] * cols)
            print("%-10s L=%5d cols=%7d  %7.3f ms  %6.1f GB/s  %.3f of peak" % (name, n, cols, ms, nbytes / ms / 1e6, nbytes / ms / 1e6 / peak), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "worker":
        worker()
    else:
        for flag in ("0", "1"):
            print("== FMB_NO_FAST1=%s" % flag, flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "worker"], env=dict(os.environ, FMB_NO_FAST1=flag), check=False)
