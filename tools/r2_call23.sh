#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
for occ in 0 7 8; do echo "V32T=0 OCC=$occ"; FMB_V32T=0 FMB_V32_OCC=$occ timeout 60 build/cbench $L circ 256; FMB_V32T=0 FMB_V32P=0 FMB_V32_OCC=$occ timeout 60 build/cbench $L fourier 256; done
for cv in 50 60 75; do echo "V32T=0 OCC=7 CARVEOUT=$cv"; FMB_V32T=0 FMB_V32_OCC=7 FMB_V32_CARVEOUT=$cv timeout 60 build/cbench $L circ 256; done
for ns in 2 4; do echo "V32T=0 OCC=7 STREAMS=$ns"; FMB_V32T=0 FMB_V32_OCC=7 FMB_PIPE_STREAMS=$ns timeout 60 build/cbench $L circ 256; done
echo "V32T=0 OCC=7 MB=24"; FMB_V32T=0 FMB_V32_OCC=7 FMB_PIPE_MB=24 timeout 60 build/cbench $L circ 256
echo "V32T=0 OCC=7 1024 cols"; FMB_V32T=0 FMB_V32_OCC=7 timeout 100 build/cbench $L circ 1024 5 2
} > gpurun_out/c23.txt 2>&1
cat gpurun_out/c23.txt
