#!/bin/bash
for cfg in "3 16" "4 16"; do
  set -- $cfg
  FMB_PIPE_STREAMS=$1 FMB_PIPE_MB=$2 python tools/sweep.py 256 2>&1 | tail -1 | sed -e 's/chk [^|]*|/|/'
done
