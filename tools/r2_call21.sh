#!/bin/bash
mkdir -p gpurun_out
{
for lib in build/alt/lib_noall.so build/alt/lib_noldst.so; do
for occ in 0 1; do for m in 0 3 4; do echo "$lib OCC=$occ MSHAPE=$m"; FMB_V32T=0 FMB_V32_OCC=$occ FMB_V32_MSHAPE=$m timeout 60 build/cbench $lib circ 256; done; done
FMB_V32T=0 FMB_V32P=0 timeout 60 build/cbench $lib fourier 256; FMB_V32T=0 FMB_V32P=0 FMB_V32_OCC=1 timeout 60 build/cbench $lib fourier 256
done
} > gpurun_out/c21.txt 2>&1
cat gpurun_out/c21.txt
