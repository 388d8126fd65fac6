#!/usr/bin/env bash
# FWHT: slab size x stream count sweep with the 64-value first pass
LIB=fastmat_b200/lib/libfastmat_b200.so
for mb in 8 12 16 20 24 32; do for ns in 2 3 4 6; do
  echo -n "MB=$mb NS=$ns  "; FMB_FWHT_PIPE_MB=$mb FMB_FWHT_PIPE_STREAMS=$ns timeout 120 build/cbench $LIB had 4096 5 | tail -1
done; done
