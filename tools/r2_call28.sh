#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
for op in circ fourier toep kron; do build/cbench $L $op 1024 5 2; done
build/cbench $L blue 256 5; build/cbench $L f16 64 20
} > gpurun_out/c28.txt 2>&1
cat gpurun_out/c28.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
