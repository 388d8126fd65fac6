#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
G=build/alt/lib_twg.so
T=build/alt/lib_timing.so
{
echo "== stage twiddles in shared memory + hoisted chain (main) vs through L1 (lib_twg)"
for lib in $G $L; do for op in circ fourier toep kron; do FMB_V32P=0 build/cbench $lib $op 256; done; done
echo "== phase timing (main design)"
build/cbench $T circ 256 3
FMB_V32P=0 build/cbench $T fourier 256 3
echo "== shapes with the new kernels"
for occ in 0 1; do for msh in 0 2 3; do echo "OCC=$occ MSHAPE=$msh"; FMB_V32_OCC=$occ FMB_V32_MSHAPE=$msh build/cbench $L circ 256; done; done
echo "== 1024 columns"
build/cbench $L circ 1024 5 2
build/cbench $L toep 1024 5 2
} > gpurun_out/c5.txt 2>&1
cat gpurun_out/c5.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.txt 2>&1; tail -3 gpurun_out/c5_pytest.txt
