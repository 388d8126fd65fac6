#!/bin/bash
mkdir -p gpurun_out
{
  FMB_V32P=0 timeout 200 python tools/sweep_v32p.py fourier 1024
  for pr in 0 1; do FMB_V32P_PAIR=$pr timeout 200 python tools/sweep_v32p.py fourier 1024; done
  for pr in 0 1; do FMB_V32P_PAIR=$pr FMB_V32P_PROMO=2 timeout 200 python tools/sweep_v32p.py fourier 1024; done
  for pr in 0 1; do FMB_V32P_PAIR=$pr timeout 200 python tools/sweep_v32p.py kron 1024; done
  for pr in 0 1; do FMB_V32P=2 FMB_V32P_PAIR=$pr timeout 200 python tools/sweep_v32p.py circulant 1024; done
} 2>&1 | grep -v -i warn | tee gpurun_out/sweep_v32p.log
FMB_V32P_PAIR=1 timeout 600 python -m pytest tests -m gpu -x -q -k "v32p" > gpurun_out/pytest_v32p.log 2>&1; tail -2 gpurun_out/pytest_v32p.log
