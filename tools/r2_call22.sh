#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
for pad in 0 8192 16384 24576 32768 40000; do echo "OCC=0 V32T=0 SMEM_PAD=$pad (L1 = 228 KB - 2 x (67.7 KB + pad))"; FMB_V32T=0 FMB_V32_SMEM_PAD=$pad timeout 60 build/cbench $L circ 256; done
} > gpurun_out/c22.txt 2>&1
cat gpurun_out/c22.txt
