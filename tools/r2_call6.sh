#!/bin/bash
mkdir -p gpurun_out
L=fastmat_b200/lib/libfastmat_b200.so
{
echo "== L2 prefetch of a later slab's input by the first pass (FMB_V32_PF = slabs ahead)"
for pf in 0 1 2 3 6; do echo "PF=$pf"; FMB_V32_PF=$pf build/cbench $L circ 256; done
for pf in 0 3; do echo "fourier per-pass PF=$pf"; FMB_V32P=0 FMB_V32_PF=$pf build/cbench $L fourier 256; done
} > gpurun_out/c6.txt 2>&1
cat gpurun_out/c6.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c6_pytest.txt 2>&1; tail -5 gpurun_out/c6_pytest.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; tail -c 600 gpurun_out/c6_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c6_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sustained','e2e','output_check','cpu_baseline','clocks')})
print(d['roofline']['frac'], d['roofline']['note'][-200:])
for k,v in d['extras'].items(): print(k, v)
PY
