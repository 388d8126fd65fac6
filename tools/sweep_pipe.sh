#!/bin/bash
for cfg in "0 3 16" "1 3 16" "0 1 16" "0 4 16" "0 3 32"; do
  set -- $cfg
  echo "no_v32=$1"; FMB_NO_V32=$1 FMB_PIPE_STREAMS=$2 FMB_PIPE_MB=$3 python tools/sweep.py 256 2>&1 | tail -1
done
