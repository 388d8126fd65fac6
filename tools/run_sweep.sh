#!/bin/bash
mkdir -p gpurun_out
{
  FMB_V32P=0 timeout 200 python tools/sweep_v32p.py circulant 1024
  FMB_V32P=2 timeout 200 python tools/sweep_v32p.py circulant 1024
  FMB_V32P=2 FMB_V32P_MLDG=1 timeout 200 python tools/sweep_v32p.py circulant 1024
  FMB_V32P=2 FMB_V32P_MLDG=1 FMB_V32P_AHEAD=3 timeout 200 python tools/sweep_v32p.py circulant 1024
  FMB_V32P=2 FMB_V32P_MLDG=1 FMB_V32P_SLAB=1 timeout 200 python tools/sweep_v32p.py circulant 1024
  FMB_V32P=2 FMB_V32P_MLDG=1 FMB_V32P_SLAB=4 timeout 200 python tools/sweep_v32p.py circulant 1024
  FMB_V32P=2 FMB_V32P_MLDG=1 FMB_V32P_MIX=1 timeout 200 python tools/sweep_v32p.py circulant 1024
} 2>&1 | grep -v -i warn | tee gpurun_out/sweep_v32p.log
CHECK_COLS=64 CHECK_ONLY=defaults FMB_V32P_MLDG=1 timeout 300 python tools/check_v32p.py 2>&1 | grep -E "^==|MISMATCH|CHECK|identical|TIMEOUT|FAILED"
