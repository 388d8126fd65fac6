#!/usr/bin/env bash
# is the pipelined-slab schedule limited by the host launch rate?  the same applies as one CUDA graph launch
LIB=fastmat_b200/lib/libfastmat_b200.so
for g in "" 1; do
for cfg in "16 3" "12 4" "8 4" "8 6" "4 8" "4 12" "6 8"; do set -- $cfg
  echo -n "GRAPH=$g MB=$1 NS=$2  "; CBENCH_GRAPH=$g FMB_FWHT_PIPE_MB=$1 FMB_FWHT_PIPE_STREAMS=$2 timeout 120 build/cbench $LIB had 4096 5 | tail -1
done; done
for g in "" 1; do
for op in circ toep; do
  echo -n "GRAPH=$g  "; CBENCH_GRAPH=$g timeout 120 build/cbench $LIB $op 1024 5 | tail -1
  echo -n "GRAPH=$g PIPE_MB=8 NS=4 "; CBENCH_GRAPH=$g FMB_PIPE_MB=8 FMB_PIPE_STREAMS=4 timeout 120 build/cbench $LIB $op 1024 5 | tail -1
done; done
