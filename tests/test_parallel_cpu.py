"""Host logic of the multi-GPU path on CPU: world_size-2 gloo process group, column sharding + gather.
(The apply itself needs a GPU; here a stand-in column-wise operator checks that shards reassemble exactly.)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_cols, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from fastmat_b200 import parallel as par
    g = torch.Generator().manual_seed(7)
    x = torch.randn((16, num_cols), generator=g, dtype=torch.float64).to(torch.complex128)
    xl = par.shard_columns(x)
    yl = torch.fft.fft(xl, dim=0)                     # stand-in for M.forward: any column-wise operator
    full = par.gather_columns(yl, num_cols)
    ok_all = torch.allclose(full, torch.fft.fft(x, dim=0))
    on0 = par.gather_columns(yl, num_cols, dst=0)
    ok_dst = (on0 is None) if rank != 0 else torch.allclose(on0, torch.fft.fft(x, dim=0))
    q.put((rank, bool(ok_all), bool(ok_dst), tuple(xl.shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize('num_cols', [8, 7])
def test_column_sharding_and_gather_gloo(num_cols):
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_cols, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res), res
    assert sum(r[3][1] for r in res) == num_cols


def test_shard_bounds_cover_and_balance():
    from fastmat_b200.parallel import shard_bounds
    for m in (0, 1, 7, 8, 1000, 1024):
        for w in (1, 2, 3, 8):
            blocks = [shard_bounds(m, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == m
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)
