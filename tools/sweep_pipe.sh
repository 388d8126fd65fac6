#!/bin/bash
for lib in "" build/alt/lib_noload.so build/alt/lib_NOSTORE.so build/alt/lib_NOTW.so build/alt/lib_NOLDST.so build/alt/lib_NOALL.so; do
  echo "lib=$lib"; FMB_LIB_PATH=${lib:+$PWD/$lib} python tools/sweep.py 256 2>&1 | tail -1 | sed -e 's/chk [^|]*|/|/'
done
