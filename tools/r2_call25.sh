#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{ for c in 2048 4096; do build/cbench $L had $c 5 2; done; FMB_FWHT_PIPE_STREAMS=1 build/cbench $L had 128 5; } > gpurun_out/c25.txt 2>&1
cat gpurun_out/c25.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
