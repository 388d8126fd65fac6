#!/usr/bin/env bash
# FWHT: L2 eviction hints and DRAM traffic of one apply under its real concurrency
LIB=fastmat_b200/lib/libfastmat_b200.so
for h in 0 1 2 3 4 8 7 15 11 9; do
  echo "== had 4096 HINT=$h"; FMB_FWHT_HINT=$h timeout 120 build/cbench $LIB had 4096 5 | tail -1
done
for h in 0 15; do
FMB_FWHT_HINT=$h CBENCH_PROFILE=1 timeout 300 ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none build/cbench $LIB had 1024 1 > gpurun_out/r2_range_had1024_h$h.txt 2>&1
tail -12 gpurun_out/r2_range_had1024_h$h.txt
done
