#!/bin/bash
# ncu capture of the fused persistent kernel (one launch = the whole operator apply)
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 1 -c 1 -o gpurun_out/$1 python tools/prof_one.py $2 $3 2>&1 | tail -1
