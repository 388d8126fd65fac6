import torch, sys
sys.path.insert(0, '/root/repo')
import fastmat_b200 as fm
for n, cols in ((8192, 8192), (2**14, 4096), (2**15, 2048), (2**16, 1024), (2**18, 256)):
    x = torch.view_as_complex(torch.randn((n, cols, 2), dtype=torch.float32, device="cuda"))
    F = fm.Fourier(n)
    ref = torch.fft.fft(x[:, :8].to(torch.complex128), dim=0)
    y = F.forward(x)
    err = float((y[:, :8] - ref).abs().max() / ref.abs().max())
    for _ in range(3): y = F.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): y = F.forward(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("rowmajor fourier L=%d cols=%d %.3f ms %.3f of peak err %.2e" % (n, cols, ms, 16.0 * n * cols / ms / 1e6 / 6449.7, err), flush=True)
