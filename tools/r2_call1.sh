#!/bin/bash
# round-2 GPU call 1: micro-benchmarks (FP32 packed issue, L2 bandwidth), kernel-variant sweep, parity of the new defaults
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 120 build/ubench_fp32 > gpurun_out/c1_ubench_fp32.txt 2>&1
timeout 200 build/ubench_l2 > gpurun_out/c1_ubench_l2.txt 2>&1
timeout 900 python tools/sweep_cbench.py kernels 256 > gpurun_out/c1_sweep.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.txt 2>&1
tail -3 gpurun_out/c1_pytest.txt
FMB_V32_OCC=1 FMB_V32_MSHAPE=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "circulant or toeplitz or fourier or kron" > gpurun_out/c1_pytest_alt.txt 2>&1
tail -3 gpurun_out/c1_pytest_alt.txt
cat gpurun_out/c1_ubench_fp32.txt; cat gpurun_out/c1_ubench_l2.txt; cat gpurun_out/c1_sweep.txt
