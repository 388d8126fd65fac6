#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c10_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/c10_pytest.txt 2>&1; tail -5 gpurun_out/c10_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c10_bench2.json 2> gpurun_out/c10_bench2.err
tail -c 1500 gpurun_out/c10_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c10_bench2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','sustained','output_check','config5_solvers','clocks')})
print(d['e2e'])
PY
head -12 gpurun_out/c10_topo.txt
