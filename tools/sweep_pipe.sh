#!/bin/bash
for lib in build/alt/lib_BASE.so build/alt/lib_MID.so build/alt/lib_ST.so build/alt/lib_MIDST.so build/alt/lib_BASE.so; do
  echo "lib=$lib"; FMB_LIB_PATH=${lib:+$PWD/$lib} python tools/sweep.py 256 2>&1 | tail -1 | sed -e 's/chk [^|]*|/|/'
done
