#!/bin/bash
for cfg in "2 16" "2 24" "3 8" "3 24" "4 8" "4 24" "5 16" "6 16" "6 8"; do
  set -- $cfg
  FMB_PIPE_STREAMS=$1 FMB_PIPE_MB=$2 python tools/sweep.py 256 2>&1 | tail -1 | sed -e 's/chk [^|]*|/|/'
done
