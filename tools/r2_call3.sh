#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
P=build/alt/lib_twp.so
{
echo "== main lib reference points"
build/cbench $L circ 256
FMB_V32P=2 build/cbench $L circ 256
FMB_V32P=0 build/cbench $L fourier 256
echo "== producer-side stage twiddle (lib_twp)"
for occ in 0 1; do for msh in 0 2 3 4; do echo "OCC=$occ MSHAPE=$msh"; FMB_V32_OCC=$occ FMB_V32_MSHAPE=$msh build/cbench $P circ 256; done; done
for occ in 0 1; do echo "fourier per-pass OCC=$occ"; FMB_V32P=0 FMB_V32_OCC=$occ build/cbench $P fourier 256; done
echo "== OCC=1 with bigger slabs / more streams (lib_twp)"
for mb in 16 24 32; do for ns in 3 4 6; do echo "OCC=1 MSHAPE=3 PIPE_MB=$mb STREAMS=$ns"; FMB_V32_OCC=1 FMB_V32_MSHAPE=3 FMB_PIPE_MB=$mb FMB_PIPE_STREAMS=$ns build/cbench $P circ 256; done; done
} > gpurun_out/c3.txt 2>&1
cat gpurun_out/c3.txt
