#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
for rep in 1 2; do for t in 0 1 2; do echo "V32T=$t (rep $rep)"; FMB_V32T=$t timeout 100 build/cbench $L circ 1024 5 2; FMB_V32T=$t timeout 100 build/cbench $L toep 1024 5 1; done; done
} > gpurun_out/c15.txt 2>&1
cat gpurun_out/c15.txt
