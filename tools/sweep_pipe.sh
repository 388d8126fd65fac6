#!/bin/bash
for cfg in "1 16" "3 16" "3 32" "4 16"; do
  set -- $cfg
  echo "tile=13"; FMB_PIPE_STREAMS=$1 FMB_PIPE_MB=$2 python tools/sweep.py 256 2>&1 | tail -1
  echo "tile=12"; FMB_LIB_PATH=$PWD/build/alt/lib_tile12.so FMB_PIPE_STREAMS=$1 FMB_PIPE_MB=$2 python tools/sweep.py 256 2>&1 | tail -1
done
