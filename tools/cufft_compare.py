"""Library comparator: cuFFT through torch.fft on the same shapes, for context next to the
hand-written kernels.  Not part of the product path; prints ms per 256 / 1024 columns."""
import sys
import torch

N = 1 << 20
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.view_as_complex(torch.randn((cols, N, 2), device='cuda'))            # one contiguous signal per row = column-major (N, cols)
s = torch.view_as_complex(torch.randn((N, 2), device='cuda'))


def timed(f, k=5):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


gb = 16.0 * N * cols / 1e9
t = timed(lambda: torch.fft.fft(x, dim=1))
print("cuFFT fft 2^20 c64, %d cols: %.3f ms  %.0f GB/s algorithmic (%.3f of 6449.7)" % (cols, t, gb / t * 1e3, gb / t * 1e3 / 6449.7))
t = timed(lambda: torch.fft.ifft(torch.fft.fft(x, dim=1) * s, dim=1))
print("cuFFT circulant (fft, multiply, ifft) %d cols: %.3f ms  %.0f GB/s algorithmic (%.3f of 6449.7)" % (cols, t, gb / t * 1e3, gb / t * 1e3 / 6449.7))
