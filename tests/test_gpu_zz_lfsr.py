"""(Named zz so that it runs after the parity suite of the headline path.)  GPU parity of LFSRCirculant: class layer -> C-ABI (scatter, FWHT, gather kernels) against the
fixtures frozen from the real reference (tests/golden/golden_lfsr.npz) and the numpy oracle; all bit-exact (integer
and small-integer-valued float inputs), plus the size-independent circulant properties on a 2^20-1 register."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from test_gpu_parity import fm, orc, dev, host          # noqa: F401  (fixtures + helpers)

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_lfsr.npz'))
REGISTERS = [tuple(int(v) for v in r) for r in G['registers']]


def tag(reg):
    return '%x_%x' % reg


@pytest.mark.parametrize('reg', REGISTERS, ids=tag)
def test_lfsr_circulant_golden_bit_exact(fm, reg):      # noqa: F811
    t = tag(reg)
    L = fm.LFSRCirculant(*reg)
    n = int(G[t + '_period'])
    assert L.shape == (n, n) and L.period == n and L.dtype == np.int8
    assert np.array_equal(L.states, G[t + '_states']) and np.array_equal(L.vecC, G[t + '_vecC'])
    for dt in ('int8', 'int32', 'int64', 'float64'):
        x = G['%s_%s_x' % (t, dt)]
        for lay in ('F', 'C'):
            xd = dev(x, lay)
            keep = xd.clone()
            for fn, key in ((L.forward, 'fwd'), (L.backward, 'bwd')):
                y = host(fn(xd))
                ref = G['%s_%s_%s' % (t, dt, key)]
                assert y.dtype == ref.dtype and np.array_equal(y, ref), (dt, lay, key)
            assert torch.equal(xd, keep)                               # the input is never modified
        assert np.array_equal(host(L.forward(dev(x[:, 0]))), G['%s_%s_fwd' % (t, dt)][:, 0])      # 1-D in, 1-D out


def test_lfsr_circulant_dense_reference_and_helpers(fm, orc):      # noqa: F811
    L = fm.LFSRCirculant(0x19, 0xD)
    dense = orc.dense_lfsr_circulant(0x19, 0xD)
    assert np.array_equal(host(L.reference()), dense)
    assert np.array_equal(host(L.getArray()), dense)
    assert np.array_equal(host(L.getCol(3)), dense[:, 3]) and np.array_equal(host(L.getRow(5)), dense[5, :])
    assert np.allclose(host(L.colNorms), np.sqrt(15)) and np.allclose(host(L.rowNorms), np.sqrt(15))
    x = np.arange(15 * 2, dtype=np.int32).reshape(15, 2)
    assert np.array_equal(host(L.H.forward(dev(x))), dense.T.astype(np.int32).dot(x))
    for args in ((0x19, 0), (1, 1)):
        with pytest.raises(ValueError):
            fm.LFSRCirculant(*args)
    with pytest.raises(ValueError):
        L.forward(dev(np.zeros((14, 2), dtype=np.int32)))


def test_lfsr_circulant_order20_properties(fm):        # noqa: F811
    # x^20 + x^3 + 1 is primitive: period 2^20 - 1, the embedded Hadamard is BASELINE config 3's order 20
    L = fm.LFSRCirculant((1 << 20) | (1 << 3) | 1, 1)
    n = 2 ** 20 - 1
    assert L.shape == (n, n) and L.order == 20
    c = torch.from_numpy(L.vecC.astype(np.int32)).cuda()
    e = torch.zeros((n, 2), dtype=torch.int32, device='cuda')
    e[0, 0] = 1
    e[7, 1] = 1
    y = L.forward(e)                                     # columns 0 and 7 of the circulant = the rolled sequence
    assert torch.equal(y[:, 0], c) and torch.equal(y[:, 1], torch.roll(c, 7))
    g = torch.Generator(device='cuda').manual_seed(11)
    x = torch.randint(-4, 5, (3, n), dtype=torch.int32, device='cuda', generator=g).t()
    # m-sequence autocorrelation: L^H L = (N + 1) I - 1 1^T, exact in integers
    z = L.backward(L.forward(x))
    assert torch.equal(z, (n + 1) * x - x.sum(dim=0, keepdim=True))
    # against an FFT-free definition on a few rows: y[i] = sum_j c[(i - j) mod N] x[j]
    yx = L.forward(x)
    for i in (0, 1, 12345, n - 1):
        row = c[(i - torch.arange(n, device='cuda')) % n]
        assert torch.equal(yx[i], (row[:, None] * x).sum(dim=0).to(torch.int32))


def test_partial_backward_scatter_holes_and_repeats(fm):        # noqa: F811
    """The backward of a row selection (y = 0; y[idx[r]] = x[r], fastmat/Partial.pyx:282-294) runs as a gather through
    the inverse index: rows nothing lands in are zero, a repeated index resolves like numpy (last occurrence wins)."""
    rng = np.random.default_rng(5)
    n = 1000
    for idx in (rng.permutation(n)[:333], np.array([7, 3, 7, 999, 0, 3, 7]), np.arange(n)[::-1].copy()):
        P = fm.Partial(fm.Eye(n), rows=idx)
        for dt in ('int8', 'int32', 'float64', 'complex128'):
            x = rng.integers(-100, 100, size=(idx.size, 5)).astype(dt)
            ref = np.zeros((n, 5), dtype=dt)
            ref[idx] = x
            for lay in ('F', 'C'):
                assert np.array_equal(host(P.backward(dev(x, lay))), ref), (dt, lay)
            assert np.array_equal(host(P.backward(dev(x[:, 0]))), ref[:, 0])


_SCATTER_FALLBACK = r'''
import sys, numpy as np, torch
sys.path.insert(0, '.')
import fastmat_b200 as fm
rng = np.random.default_rng(5)
n = 1000
for idx in (np.array([7, 3, 7, 999, 0, 3, 7]), rng.integers(0, n, size=400), rng.permutation(n)[:333]):
    P = fm.Partial(fm.Eye(n), rows=idx)
    for dt in ('int8', 'float64', 'complex128'):
        x = rng.integers(-100, 100, size=(idx.size, 5)).astype(dt)
        ref = np.zeros((n, 5), dtype=dt)
        ref[idx] = x
        for _ in range(3):                                   # a race would not show every time
            got = P.backward(torch.from_numpy(np.ascontiguousarray(x.T)).cuda().t()).cpu().numpy()
            assert np.array_equal(got, ref), dt
print('SCATTER_OK')
'''


def test_partial_backward_zero_scatter_path_is_deterministic_for_repeated_indices():
    """Without the inverse table (FMB_PARTIAL_INVERSE=0, or a few rows out of a huge matrix) the backward is zero-fill +
    scatter; repeated indices are reduced to their last occurrence at plan creation, so the result equals numpy's
    y[idx] = x there as well instead of depending on which thread wrote last."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FMB_PARTIAL_INVERSE='0')
    r = subprocess.run([sys.executable, '-c', _SCATTER_FALLBACK], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'SCATTER_OK' in r.stdout, r.stdout + r.stderr
