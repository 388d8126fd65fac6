#!/bin/bash
# round-2 profile artefacts: ncu launch list of the bench command (per-kernel share of a step, DRAM bytes), full captures of
# the passes of Circulant / Fourier 2^20 and of the FWHT, range-replay traffic of one apply under its real concurrency
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -s 4700 -c 1600 --csv --log-file gpurun_out/r2_launches_bench_circulant.csv python bench.py --steps 2 --warmup 3 --quick --no-e2e --no-cpu --sustain 0 > gpurun_out/r2_prof_bench.log 2>&1
tail -2 gpurun_out/r2_prof_bench.log | cut -c1-300
FMB_PIPE_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:v32 -s 9 -c 3 -f -o gpurun_out/r2_circ_slab64 build/cbench $L circ 64 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v32p -s 3 -c 1 -f -o gpurun_out/r2_fourier_v32p build/cbench $L fourier 64 1 > /dev/null 2>&1
FMB_FWHT_PIPE_STREAMS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fwht -s 6 -c 2 -f -o gpurun_out/r2_hadamard build/cbench $L had 128 1 > /dev/null 2>&1
FMB_V32T=0 CBENCH_PROFILE=1 timeout 300 ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none build/cbench $L circ 1024 1 > gpurun_out/r2_range_circ1024.txt 2>&1
tail -14 gpurun_out/r2_range_circ1024.txt
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_bench_circulant.csv
