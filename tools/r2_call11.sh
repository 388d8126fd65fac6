#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
echo "== tiles per CTA (OCC / MSHAPE: 0 = one tile, 5 = two, 6 = four)"
for occ in 0 5 6; do for msh in 0 5 6; do echo "OCC=$occ MSHAPE=$msh"; FMB_V32_OCC=$occ FMB_V32_MSHAPE=$msh build/cbench $L circ 256; done; done
for mb in 32 64; do for ns in 2 3 4; do echo "OCC=5 MSHAPE=5 PIPE_MB=$mb STREAMS=$ns"; FMB_V32_OCC=5 FMB_V32_MSHAPE=5 FMB_PIPE_MB=$mb FMB_PIPE_STREAMS=$ns build/cbench $L circ 256; done; done
echo "== Toeplitz: pruned butterflies"
for pr in 0 1; do echo "PRUNE=$pr"; FMB_V32_PRUNE=$pr build/cbench $L toep 256; FMB_V32_PRUNE=$pr build/cbench $L toepb 256; done
} > gpurun_out/c11.txt 2>&1
cat gpurun_out/c11.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c11_pytest.txt 2>&1; tail -5 gpurun_out/c11_pytest.txt
python - <<'PY' > gpurun_out/c11_rm.txt 2>&1
import sys, torch, numpy as np
sys.path.insert(0, '.')
import fastmat_b200 as fm
N = 1 << 20
rng = np.random.default_rng(0)
c = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
C, F = fm.Circulant(c), fm.Fourier(N)
x = torch.view_as_complex(torch.randn((N, 1024, 2), device='cuda'))      # row-major (N, 1024)
def timed(f, k=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print("row-major circulant 2^20 x 1024: %.2f ms" % timed(lambda: C.forward(x)))
print("row-major fourier   2^20 x 1024: %.2f ms" % timed(lambda: F.forward(x)))
xc = x.t().contiguous().t()
yr, yc = C.forward(x), C.forward(xc)
print("row-major == column-major result:", float((yr - yc).abs().max()), "is_row_major out:", yr.stride())
PY
cat gpurun_out/c11_rm.txt
