#!/bin/bash
# experiments: schedule / split knobs
for cfg in "1 16 0" "3 16 0" "3 16 8" "3 16 12" "3 16 9"; do
  set -- $cfg
  echo "split=$3 inner=row"; FMB_PIPE_STREAMS=$1 FMB_PIPE_MB=$2 FMB_SPLIT_LOG1=$3 python tools/sweep.py 256 2>&1 | tail -1
  echo "split=$3 inner=line"; FMB_LIB_PATH=$PWD/build/alt/lib_innerT.so FMB_PIPE_STREAMS=$1 FMB_PIPE_MB=$2 FMB_SPLIT_LOG1=$3 python tools/sweep.py 256 2>&1 | tail -1
done
