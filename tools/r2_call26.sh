#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 800 gpurun_out/bench_r2.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm_r2.json 2> gpurun_out/bench_ref_r2.err; tail -c 300 gpurun_out/bench_ref_r2.err
python tools/cufft_compare.py 256 > gpurun_out/cufft_r2.txt 2>&1; cat gpurun_out/cufft_r2.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2.json'))
print({k:d[k] for k in ('value','ms_per_step','sustained','output_check','config5_solvers','clocks','gpu_launches','cpu_baseline')})
print(d['e2e'])
print(d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['dominant_kernel'], d['roofline']['kernels'])
for k,v in d['extras'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
r=json.load(open('gpurun_out/bench_reference_arm_r2.json')); print('reference arm', r['value'], r['cpu_baseline'])
PY
