#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
echo "V32P (default)"; build/cbench $L fourier 1024 5 2
echo "per-pass"; FMB_V32P=0 build/cbench $L fourier 1024 5 2
echo "per-pass, TMA-fed first pass"; FMB_V32P=0 FMB_V32T_PLAIN=1 build/cbench $L fourier 1024 5 2
echo "per-pass, TMA-fed first pass, backward"; FMB_V32P=0 FMB_V32T_PLAIN=1 build/cbench $L fourierb 1024 5 2
} > gpurun_out/c27.txt 2>&1
cat gpurun_out/c27.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
