#!/bin/bash
for cfg in "1 16" "2 16" "3 8" "3 16" "3 32" "4 16" "4 8" "6 8"; do
  set -- $cfg
  echo "fwht pipe=$1 mb=$2"; FMB_FWHT_PIPE_STREAMS=$1 FMB_FWHT_PIPE_MB=$2 python tools/sweep.py 512 hadamard 2>&1 | tail -2
done
