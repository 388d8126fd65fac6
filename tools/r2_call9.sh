#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_algorithms.py -m gpu -x -q > gpurun_out/c9_pytest.txt 2>&1; tail -5 gpurun_out/c9_pytest.txt
timeout 600 python tools/omp_profile.py 128 > gpurun_out/c9_omp.txt 2>&1; cat gpurun_out/c9_omp.txt | cut -c1-200
