#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
for t in 1 2; do for o in 0 1; do for m in 0 3; do echo "V32T=$t V32T_OCC=$o MSHAPE=$m"; FMB_V32T=$t FMB_V32T_OCC=$o FMB_V32_MSHAPE=$m timeout 60 build/cbench $L circ 256; done; done; done
echo "V32T=1 V32T_OCC=1 OCC=1 MSHAPE=3"; FMB_V32T=1 FMB_V32T_OCC=1 FMB_V32_OCC=1 FMB_V32_MSHAPE=3 timeout 60 build/cbench $L circ 256
for ns in 3 4 6; do echo "V32T=1 V32T_OCC=1 STREAMS=$ns MB=24"; FMB_V32T=1 FMB_V32T_OCC=1 FMB_PIPE_STREAMS=$ns FMB_PIPE_MB=24 timeout 60 build/cbench $L circ 256; done
} > gpurun_out/c19.txt 2>&1
cat gpurun_out/c19.txt
