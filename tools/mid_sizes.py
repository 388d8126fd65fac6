import torch, sys, os
sys.path.insert(0, '/root/repo')
import fastmat_b200 as fm
for n, cols, dt in ((2**13, 8192, torch.complex64), (2**14, 4096, torch.complex64), (2**16, 1024, torch.complex64), (2**18, 256, torch.complex64), (2**16, 64, torch.complex128), (2**16, 1024, torch.complex128), (2**22, 64, torch.complex64)):
    x = torch.view_as_complex(torch.randn((cols, n, 2), dtype=torch.float64 if dt == torch.complex128 else torch.float32, device="cuda")).t()
    F = fm.Fourier(n)
    ref = torch.fft.fft(x[:, :4].to(torch.complex128), dim=0)
    y = F.forward(x)
    err = float((y[:, :4] - ref).abs().max() / ref.abs().max())
    for _ in range(3): y = F.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): y = F.forward(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    eb = 16 if dt == torch.complex128 else 8
    print("fourier L=%d cols=%d %s %.4f ms %.3f of peak err %.2e" % (n, cols, str(dt)[6:], ms, 2.0 * eb * n * cols / ms / 1e6 / 6449.7, err), flush=True)
