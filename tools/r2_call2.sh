#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
echo "== phase timing (instrumented build, 256 columns)"
build/cbench build/alt/lib_timing.so circ 256 3
FMB_V32P=0 build/cbench build/alt/lib_timing.so fourier 256 3
build/cbench build/alt/lib_timing.so toep 256 3
echo "== phase timing, single stream, 64-column slabs (FMB_PIPE_STREAMS=1)"
FMB_PIPE_STREAMS=1 FMB_SLAB_MB=512 build/cbench build/alt/lib_timing.so circ 256 3
echo "== carve-out experiments"
for cv in 0 50 60 75 100; do echo "carveout $cv"; FMB_V32_CARVEOUT=$cv build/cbench $L circ 256; done
for cv in 75 90 100; do echo "OCC=1 carveout $cv"; FMB_V32_OCC=1 FMB_V32_CARVEOUT=$cv build/cbench $L circ 256; done
echo "== 1024 columns, current defaults"
build/cbench $L circ 1024 5 2
build/cbench $L fourier 1024 5 2
build/cbench $L toep 1024 5 2
build/cbench $L kron 1024 5 2
build/cbench $L had 2048 5 2
build/cbench $L blue 256 5
build/cbench $L f16 64 20
echo "== cuFFT comparator"
python tools/cufft_compare.py 256
} > gpurun_out/c2.txt 2>&1
cat gpurun_out/c2.txt
