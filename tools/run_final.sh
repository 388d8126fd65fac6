#!/bin/bash
# Round-end session on one B200: GPU test-suite, bench (both arms), ncu captures of the persistent kernels, memcheck.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:v32p_kernel -c 1 -o gpurun_out/r1_v32p_fourier -f python tools/prof_one.py fourier 64 > gpurun_out/ncu_v32p.log 2>&1; tail -1 gpurun_out/ncu_v32p.log
FMB_V32P=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:v32p_kernel -c 1 -o gpurun_out/r1_v32p_circulant -f python tools/prof_one.py circulant 64 > gpurun_out/ncu_v32p_c.log 2>&1; tail -1 gpurun_out/ncu_v32p_c.log
FMB_V32P=2 timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py > gpurun_out/memcheck_v32p.log 2>&1; tail -2 gpurun_out/memcheck_v32p.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
