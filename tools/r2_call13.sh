#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
for t in 0 1 2 3; do echo "V32T=$t"; FMB_V32T=$t timeout 60 build/cbench $L circ 256; FMB_V32T=$t timeout 60 build/cbench $L toep 256;  FMB_V32T=$t timeout 60 build/cbench $L circb 256; done
echo "== with PIPE variants"
for ns in 2 3 4; do echo "V32T=1 STREAMS=$ns"; FMB_V32T=1 FMB_PIPE_STREAMS=$ns timeout 60 build/cbench $L circ 256; done
FMB_V32T=1 timeout 100 build/cbench $L circ 1024 5 2
} > gpurun_out/c13.txt 2>&1
cat gpurun_out/c13.txt
FMB_V32T=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "circulant or toeplitz" > gpurun_out/c13_pytest.txt 2>&1; tail -3 gpurun_out/c13_pytest.txt
