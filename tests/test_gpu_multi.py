"""GPU, world size 2 over NCCL (skipped with fewer than two visible GPUs): column-sharded solves and applies through
fastmat_b200.parallel, assembled with one all_gather / gather, checked against the reference's fixtures (the same host
logic runs under gloo on CPU in tests/test_parallel_cpu.py)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_solve_sharded_and_gather_two_ranks_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tests', '_multi_worker.py')]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and 'MULTI_OK world=2' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
