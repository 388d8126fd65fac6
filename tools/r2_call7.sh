#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
for pf in 0 1 3; do echo "PF=$pf"; FMB_V32_PF=$pf build/cbench $L circ 256; done
build/cbench $L circ 1024 5 2
build/cbench $L toep 1024 5 2
} > gpurun_out/c7.txt 2>&1
cat gpurun_out/c7.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c7_pytest.txt 2>&1; tail -3 gpurun_out/c7_pytest.txt
