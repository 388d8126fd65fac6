#!/usr/bin/env bash
LIB=fastmat_b200/lib/libfastmat_b200.so
echo -n "launch-based      "; timeout 60 build/cbench $LIB had 4096 5 | tail -1
echo -n "persist NT=128 64c"; FMB_FWHT_PERSIST=1 FMB_FWHT_PERSIST_NT=128 timeout 60 build/cbench $LIB had 64 3 | tail -1
for cfg in "128 2 2 5" "128 1 4 5" "128 1 3 5" "128 2 1 5" "128 1 2 5" "128 2 2 4" "128 2 2 3" "128 1 4 4" "128 1 4 3" "128 1 6 5" "128 4 1 5" "256 1 4 2" "256 2 2 2"; do set -- $cfg
  echo -n "persist NT=$1 SLAB=$2 DIST=$3 CTAS=$4  "; FMB_FWHT_PERSIST=1 FMB_FWHT_PERSIST_NT=$1 FMB_FWHT_PERSIST_SLAB=$2 FMB_FWHT_PERSIST_DIST=$3 FMB_FWHT_PERSIST_CTAS=$4 timeout 60 build/cbench $LIB had 4096 5 | tail -1
done
