#!/bin/bash
mkdir -p gpurun_out
T=build/alt/lib_timing.so
{
for occ in 0 1; do for msh in 0 4; do echo "== OCC=$occ MSHAPE=$msh"; FMB_V32_OCC=$occ FMB_V32_MSHAPE=$msh build/cbench $T circ 256 3; done; done
} > gpurun_out/c4.txt 2>&1
cat gpurun_out/c4.txt
