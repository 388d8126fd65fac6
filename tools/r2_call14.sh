#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
echo "== L2 persisting window on the ring"
for p in 0 1 2; do echo "L2_PERSIST=$p"; FMB_L2_PERSIST=$p timeout 60 build/cbench $L circ 256; FMB_L2_PERSIST=$p timeout 100 build/cbench $L circ 1024 5; done
FMB_L2_PERSIST=1 FMB_V32T=1 timeout 100 build/cbench $L circ 1024 5
echo "== DRAM traffic of ONE apply with its real concurrency (ncu range replay)"
CBENCH_PROFILE=1 timeout 300 ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none build/cbench $L circ 256 1 2>&1 | tail -25
FMB_L2_PERSIST=1 CBENCH_PROFILE=1 timeout 300 ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none build/cbench $L circ 256 1 2>&1 | tail -15
} > gpurun_out/c14.txt 2>&1
cat gpurun_out/c14.txt
