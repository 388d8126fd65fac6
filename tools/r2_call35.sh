#!/usr/bin/env bash
LIB=fastmat_b200/lib/libfastmat_b200.so
for op in fourier kron; do
for e in "X=0" "FMB_V32P_DELAY=2" "FMB_V32P_DELAY=3" "FMB_V32P_MIX=1" "FMB_V32P_MIX=1 FMB_V32P_DELAY=2" "FMB_V32P_AHEAD=3" "FMB_V32P_AHEAD=8" "FMB_V32P_SLAB=2" "FMB_V32P_SLAB=2 FMB_V32P_MIX=1" "FMB_V32P_AHEAD=2" "FMB_V32P_AHEAD=1"; do
  echo -n "$op $e  "; env $e timeout 120 build/cbench $LIB $op 1024 5 | tail -1
done; done
