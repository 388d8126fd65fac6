#!/bin/bash
# experiment: is the host-buffer (e2e) number sensitive to the schedule knobs or to the box?
nvidia-smi topo -m 2>/dev/null | head -8; nproc; numactl -H 2>/dev/null | head -5
for cfg in "3 0" "1 0" "3 1"; do
  set -- $cfg
  echo "PIPE=$1 NO_V32=$2"; FMB_PIPE_STREAMS=$1 FMB_NO_V32=$2 python bench.py --quick --no-cpu --steps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['value'])"
done
