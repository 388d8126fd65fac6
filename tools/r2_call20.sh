#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
FMB_PIPE_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:v32 -s 9 -c 3 -f -o gpurun_out/r2_circ_slab64 build/cbench $L circ 64 1 > gpurun_out/c20_ncu1.log 2>&1
tail -3 gpurun_out/c20_ncu1.log
# pipelined mode: kernels of 2-column slabs (256 CTAs each), profiled one at a time
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v32 -s 36 -c 3 -f -o gpurun_out/r2_circ_pipe build/cbench $L circ 24 1 > gpurun_out/c20_ncu2.log 2>&1
tail -3 gpurun_out/c20_ncu2.log
ls -la gpurun_out/*.ncu-rep
