#!/usr/bin/env bash
# FWHT: the 12-bit first pass with 64 values per thread (fwht_first12_kernel) against the three-sub-stage one
LIB=fastmat_b200/lib/libfastmat_b200.so
timeout 600 python -m pytest tests -x -q -m gpu -k "hadamard or Hadamard or fwht or lfsr" 2>&1 | tail -4
for e in "" "FMB_FWHT_NO12=1" "FMB_FWHT_PIPE_STREAMS=1" "FMB_FWHT_PIPE_STREAMS=1 FMB_FWHT_NO12=1" "FMB_FWHT_PIPE_MB=8" "FMB_FWHT_PIPE_MB=32" "FMB_FWHT_PIPE_STREAMS=4" "FMB_FWHT_PIPE_STREAMS=2"; do
  echo "== had 4096 $e"; env $e timeout 120 build/cbench $LIB had 4096 5 | tail -1
done
