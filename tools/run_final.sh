#!/bin/bash
# Round-end session on one B200: GPU test-suite, bench (both arms), delay sweep of the persistent Fourier path, memcheck.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.json
{
  for d in 2 3 4 5 6; do FMB_V32P_DELAY=$d timeout 200 python tools/sweep_v32p.py fourier 512; done
  FMB_V32P=0 timeout 200 python tools/sweep_v32p.py fourier 1024
  timeout 200 python tools/sweep_v32p.py fourier 1024
} 2>&1 | grep -v -i warn | tee gpurun_out/sweep_delay.log
FMB_V32P=2 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/memcheck_v32p.log 2>&1; tail -4 gpurun_out/memcheck_v32p.log
