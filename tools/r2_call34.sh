#!/usr/bin/env bash
# persistent plain transforms: intermediate in y (in place) against the ring in the workspace
LIB=fastmat_b200/lib/libfastmat_b200.so
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fourier or kron or v32p or V32P or persistent or concurrency or switch" 2>&1 | tail -4
for op in fourier fourierb kron; do
for e in "FMB_V32P_INPLACE=0" "FMB_V32P_INPLACE=1" "FMB_V32P_INPLACE=1 FMB_V32P_HINTS=0" "FMB_V32P_INPLACE=0 FMB_V32P_HINTS=0"; do
  echo -n "$op $e  "; env $e timeout 120 build/cbench $LIB $op 1024 5 2 | tail -1
done; done
