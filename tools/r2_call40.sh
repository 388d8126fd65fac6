#!/usr/bin/env bash
# persistent order-20 FWHT (FMB_FWHT_PERSIST=1): correctness (checksum against the launch-based path), then a parameter sweep
LIB=fastmat_b200/lib/libfastmat_b200.so
echo -n "launch-based      "; timeout 60 build/cbench $LIB had 4096 5 | tail -1
echo -n "persistent 64 cols"; FMB_FWHT_PERSIST=1 timeout 60 build/cbench $LIB had 64 3 | tail -1
echo -n "launch     64 cols"; timeout 60 build/cbench $LIB had 64 3 | tail -1
for cfg in "2 2 2" "2 1 2" "2 3 2" "4 1 2" "4 2 2" "1 2 2" "1 4 2" "1 3 2" "2 2 1" "3 2 2" "8 1 2"; do set -- $cfg
  echo -n "persist SLAB=$1 DIST=$2 CTAS=$3  "; FMB_FWHT_PERSIST=1 FMB_FWHT_PERSIST_SLAB=$1 FMB_FWHT_PERSIST_DIST=$2 FMB_FWHT_PERSIST_CTAS=$3 timeout 60 build/cbench $LIB had 4096 5 | tail -1
done
