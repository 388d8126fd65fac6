#!/bin/bash
# One GPU-box session: GPU test-suite + A/B timings of the persistent (V32P) path against the per-pass kernels.
mkdir -p gpurun_out
[ -n "$SKIP_TESTS" ] || timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
cat > /tmp/kron_time.py <<'PY'
import sys, os, torch
sys.path.insert(0, '.')
import fastmat_b200 as fm
N, cols = 1 << 20, 512
x = torch.view_as_complex(torch.randn((cols, N, 2), device='cuda')).t()
def timed(f, k=5):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
K, F = fm.Kron(fm.Fourier(1024), fm.Fourier(1024)), fm.Fourier(N)
gb = 16.0 * N * cols / 1e9
for name, M in (('kron', K), ('fourier', F)):
    t = min(timed(lambda: M.forward(x)) for _ in range(3))
    print('%s V32P=%s SLAB=%s AHEAD=%s: %.3f ms %.1f%%' % (name, os.environ.get('FMB_V32P'), os.environ.get('FMB_V32P_SLAB'), os.environ.get('FMB_V32P_AHEAD'), t, gb / t * 1e3 / 6550.1 * 100))
PY
{
  FMB_V32P=0 timeout 300 python /tmp/kron_time.py
  for slab in 1 2 4; do for ahead in 2 3 5 8; do FMB_V32P=1 FMB_V32P_SLAB=$slab FMB_V32P_AHEAD=$ahead timeout 300 python /tmp/kron_time.py; done; done
} 2>&1 | grep -v Warning > gpurun_out/ab2.log
cat gpurun_out/ab2.log
