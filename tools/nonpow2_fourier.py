import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import fastmat_b200 as fm
for n, cols in ((100, 65536), (1000, 16384), (1500, 8192), (3000, 4096), (6144, 1024), (110592, 256), (5000, 1024), (786432, 32), (12288, 1024)):
    x = torch.view_as_complex(torch.randn((cols, n, 2), dtype=torch.float32, device='cuda')).t()
    F = fm.Fourier(n)
    ref = torch.fft.fft(x[:, :2].to(torch.complex128), dim=0)
    y = F.forward(x)
    err = float((y[:, :2] - ref).abs().max() / (torch.linalg.vector_norm(x[:, :2].to(torch.complex128), dim=0).max() * np.log2(n)))
    errb = float((F.backward(y)[:, :2] / n - x[:, :2]).abs().max())
    for _ in range(2): y = F.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): y = F.forward(x)
    e1.record(); torch.cuda.synchronize()
    print('fourier n=%d cols=%d inner=%d %.3f ms err %.1e roundtrip %.1e' % (n, cols, int(F._plan.info.inner_size), e0.elapsed_time(e1) / 5, err, errb), flush=True)
