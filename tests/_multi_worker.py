"""Worker of tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU, NCCL): the right-hand sides of the
reference's solver fixtures are sharded over the ranks, every rank solves its shard through parallel.solve_sharded, the
result is assembled with parallel.gather_columns (NCCL all_gather) and every rank checks it against the fixtures frozen
from the reference; a sharded plain apply is checked the same way."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    import fastmat_b200 as fm
    from fastmat_b200 import parallel as fpar
    GA = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_algorithms.npz'))
    d = GA['cs_d'].astype(np.complex128)
    A = fm.Product(fm.Partial(fm.Fourier(256), rows=GA['cs_rows']), fm.Diag(d))
    lam, steps, k = GA['cs_params']
    b = torch.from_numpy(np.ascontiguousarray(GA['cs_b'].T)).to(dev).t()
    ncols = b.shape[1]
    b_loc = fpar.shard_columns(b, rank, world)
    got = fpar.gather_columns(fpar.solve_sharded(fm.algorithms.ISTA(A, numLambda=float(lam), numMaxSteps=int(steps)), b_loc), ncols)
    ref = GA['cs_ista']
    assert got.shape == ref.shape and np.abs(got.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
    got = fpar.gather_columns(fpar.solve_sharded(fm.algorithms.OMP(A, numMaxSteps=int(k)), b_loc, share_step_size=False), ncols)
    ref = GA['cs_omp']
    assert np.array_equal(got.cpu().numpy() != 0, ref != 0)
    assert np.allclose(got.cpu().numpy(), ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    # plain sharded apply + gather to rank 0 only
    y = fpar.gather_columns(fpar.apply_sharded(A, fpar.shard_columns(torch.from_numpy(np.ascontiguousarray(GA['cs_omp'].T)).to(dev).t(), rank, world)),
                            ncols, dst=0)
    if rank == 0:
        full = A.forward(torch.from_numpy(np.ascontiguousarray(GA['cs_omp'].T)).to(dev).t())
        assert torch.equal(y, full)
    else:
        assert y is None
    dist.barrier()
    if rank == 0:
        print('MULTI_OK world=%d' % world)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
