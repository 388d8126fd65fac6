#!/bin/bash
mkdir -p gpurun_out
{
  FMB_V32P=0 timeout 200 python tools/sweep_v32p.py fourier 512
  for mix in 0 1; do for ahead in 1 2 3 5; do FMB_V32P=1 FMB_V32P_MIX=$mix FMB_V32P_AHEAD=$ahead timeout 200 python tools/sweep_v32p.py fourier 512; done; done
  FMB_V32P=0 timeout 200 python tools/sweep_v32p.py fourier 512
  FMB_V32P=0 timeout 200 python tools/sweep_v32p.py circulant 1024
  for mix in 0 1; do for ahead in 1 3 5; do FMB_V32P=2 FMB_V32P_MIX=$mix FMB_V32P_AHEAD=$ahead timeout 200 python tools/sweep_v32p.py circulant 1024; done; done
  FMB_V32P=2 FMB_V32P_MIX=0 FMB_V32P_AHEAD=3 FMB_V32P_SLAB=1 timeout 200 python tools/sweep_v32p.py circulant 1024
  FMB_V32P=2 FMB_V32P_MIX=0 FMB_V32P_AHEAD=3 FMB_V32P_SLAB=4 timeout 200 python tools/sweep_v32p.py circulant 1024
  FMB_V32P=0 timeout 200 python tools/sweep_v32p.py circulant 1024
} 2>&1 | grep -v -i warn | tee gpurun_out/sweep_v32p.log
