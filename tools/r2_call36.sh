#!/usr/bin/env bash
# Circulant on the per-pass route: intermediate in y (in place) against the ring
LIB=fastmat_b200/lib/libfastmat_b200.so
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "circulant or Circulant or switch or concurrency or big or toeplitz" 2>&1 | tail -4
for op in circ circb; do
for e in "FMB_V32_INPLACE=0" "FMB_V32_INPLACE=1" "FMB_V32_INPLACE=1 FMB_PIPE_MB=8" "FMB_V32_INPLACE=1 FMB_PIPE_MB=32" "FMB_V32_INPLACE=1 FMB_PIPE_STREAMS=4" "FMB_V32_INPLACE=1 FMB_PIPE_STREAMS=2" "FMB_V32_INPLACE=1 FMB_PIPE_MB=8 FMB_PIPE_STREAMS=6" "FMB_V32_INPLACE=1 FMB_PIPE_MB=32 FMB_PIPE_STREAMS=2" "FMB_V32_INPLACE=1 FMB_V32T=0" "FMB_V32_INPLACE=1 FMB_PIPE_MB=24"; do
  echo -n "$op $e  "; env $e timeout 120 build/cbench $LIB $op 1024 5 2 | tail -1
done; done
