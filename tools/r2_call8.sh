#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c8_pytest.txt 2>&1; tail -15 gpurun_out/c8_pytest.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err; tail -c 1500 gpurun_out/c8_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c8_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','sustained','output_check','config5_solvers')})
print(d['e2e'])
for k,v in d['extras'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
PY
