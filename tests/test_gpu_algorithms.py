"""GPU: fastmat_b200.algorithms (ISTA / FISTA / OMP on device tensors, operators applied through the C-ABI) against the
fixtures frozen from the real reference's solvers and against the numpy oracle; plus BASELINE config 5 at full operator
size (Product(Partial(Fourier(2^18)), Diag), k-sparse recovery) checked through properties."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GA = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_algorithms.npz'))


@pytest.fixture(scope='module')
def fm():
    import fastmat_b200
    assert torch.cuda.is_available()
    return fastmat_b200


def build(fm, tag, dtype):
    if tag == 'cs':
        d = GA['cs_d'].astype(dtype)
        return fm.Product(fm.Partial(fm.Fourier(256), rows=GA['cs_rows']), fm.Diag(d))
    d = GA['had_d'].astype(np.float64 if dtype == np.complex128 else np.float32)
    return fm.Product(fm.Partial(fm.Hadamard(8), rows=GA['had_rows']), fm.Diag(d))


def colmajor(a):
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()


@pytest.mark.parametrize('tag', ['cs', 'had'])
def test_largest_singular_value(fm, tag):
    A = build(fm, tag, np.complex128)
    assert abs(A.largestSingularValue - float(GA[tag + '_lsv'])) <= 1e-6 * float(GA[tag + '_lsv'])


@pytest.mark.parametrize('tag', ['cs', 'had'])
@pytest.mark.parametrize('alg', ['ISTA', 'FISTA'])
def test_ista_double_matches_reference(fm, tag, alg):
    A = build(fm, tag, np.complex128)
    lam, steps, _ = GA[tag + '_params']
    b = GA[tag + '_b']
    if tag == 'had':
        b = b.astype(np.float64)
    got = getattr(fm.algorithms, alg)(A, numLambda=float(lam), numMaxSteps=int(steps)).process(colmajor(b)).cpu().numpy()
    ref = GA['%s_%s' % (tag, alg.lower())]
    assert got.shape == ref.shape
    # step size comes from a power iteration instead of ARPACK: agreement to ~1e-6, same support
    assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
    big = np.abs(ref) > 1e-3 * np.abs(ref).max()
    assert np.array_equal((got != 0)[big], (ref != 0)[big])


def test_ista_single_precision(fm):
    A = build(fm, 'cs', np.complex64)
    lam, steps, _ = GA['cs_params']
    b = GA['cs_b'].astype(np.complex64)
    got = fm.algorithms.ISTA(A, numLambda=float(lam), numMaxSteps=int(steps)).process(colmajor(b))
    assert got.dtype == torch.complex64
    ref = GA['cs_ista']
    assert np.abs(got.cpu().numpy() - ref).max() <= 2e-4 * np.abs(ref).max()


def test_ista_numpy_in_numpy_out_and_1d(fm):
    A = build(fm, 'cs', np.complex128)
    lam, steps, _ = GA['cs_params']
    r = fm.algorithms.ISTA(A, numLambda=float(lam), numMaxSteps=int(steps)).process(GA['cs_b'][:, 0].copy())
    assert isinstance(r, np.ndarray) and r.shape == (256,)
    assert np.abs(r - GA['cs_ista'][:, 0]).max() <= 1e-5 * np.abs(GA['cs_ista']).max()


@pytest.mark.parametrize('tag', ['cs', 'had'])
def test_omp_matches_reference(fm, tag):
    A = build(fm, tag, np.complex128)
    k = int(GA[tag + '_params'][2])
    b = GA[tag + '_b']
    got = fm.algorithms.OMP(A, numMaxSteps=k).process(colmajor(b)).cpu().numpy()
    ref = GA[tag + '_omp']
    assert np.array_equal(got != 0, ref != 0)                       # identical support
    assert np.allclose(got, ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())


def test_soft_threshold_kernel(fm):
    from oracle import algorithms_oracle as ao
    rng = np.random.default_rng(5)
    for dt in (np.float32, np.float64, np.complex64, np.complex128):
        x = rng.standard_normal((1000, 7)) + (1j * rng.standard_normal((1000, 7)) if np.dtype(dt).kind == 'c' else 0)
        x = x.astype(dt)
        g = (rng.standard_normal((1000, 7)) + (1j * rng.standard_normal((1000, 7)) if np.dtype(dt).kind == 'c' else 0)).astype(dt)
        step, xn = fm.algorithms.ista_step(colmajor(x), colmajor(g), 0.3, 0.4)
        ref_step = x - dt(0.3).real * g
        ref = ao.soft_threshold(ref_step, 0.4)
        tol = 1e-5 if np.dtype(dt).itemsize in (4, 8) and dt in (np.float32, np.complex64) else 1e-12
        assert np.abs(step.cpu().numpy() - ref_step).max() <= tol * 10
        assert np.abs(xn.cpu().numpy() - ref).max() <= tol * 10


def test_errors(fm):
    A = build(fm, 'cs', np.complex128)
    with pytest.raises(TypeError):
        fm.algorithms.ISTA(np.eye(3))
    with pytest.raises(ValueError):
        fm.algorithms.OMP(A).process(GA['cs_b'])                     # numMaxSteps = 0
    with pytest.raises(ValueError):
        fm.algorithms.ISTA(A, numMaxSteps=0).process(GA['cs_b'])
    with pytest.raises(AttributeError):
        fm.algorithms.ISTA(A).process(GA['cs_b'], noSuchParameter=1)
    with pytest.raises(ValueError):
        fm.algorithms.ISTA(A).process(np.zeros((4, 4, 4)))


def test_config5_compressed_sensing_recovery(fm):
    """BASELINE config 5 at full operator size: A = Partial(Fourier(2^18), 2^16 rows) * Diag(unit modulus), k = 32."""
    n, m, k, L = 1 << 18, 1 << 16, 32, 16
    rng = np.random.default_rng(2026)
    rows = np.sort(rng.choice(n, m, replace=False))
    d = np.exp(2j * np.pi * rng.random(n)).astype(np.complex64)
    A = fm.Product(fm.Partial(fm.Fourier(n), rows=rows), fm.Diag(d))
    x = np.zeros((n, L), dtype=np.complex64)
    for c in range(L):
        idx = rng.choice(n, k, replace=False)
        x[idx, c] = (2 + rng.random(k)) * np.exp(2j * np.pi * rng.random(k))
    xd = colmajor(x)
    b = A.forward(xd)
    # all rows of F have unit-modulus entries, d too: sigma_max^2 = n exactly (rows of a scaled unitary)
    assert abs(A.largestSingularValue - np.sqrt(n)) <= 1e-4 * np.sqrt(n)
    # OMP: exact support and values
    got = fm.algorithms.OMP(A, numMaxSteps=k).process(b)
    assert got.dtype == torch.complex128
    assert torch.equal(got != 0, xd != 0)
    assert float((got - xd).abs().max()) <= 1e-3
    # ISTA: the k largest entries of every column sit on the true support, residual decreases
    res = fm.algorithms.ISTA(A, numLambda=2.0 * np.sqrt(n) * 0.0 + 50.0, numMaxSteps=30).process(b)
    top = torch.topk(res.abs(), k, dim=0).indices.cpu().numpy()
    for c in range(L):
        assert set(top[:, c]) == set(np.nonzero(x[:, c])[0])
    r0 = float(torch.linalg.vector_norm(b))
    r1 = float(torch.linalg.vector_norm(A.forward(res) - b))
    assert r1 < 0.2 * r0


def test_norm_and_singular_value_shortcuts(fm):
    """colNorms / rowNorms / largestSingularValue overrides (fastmat/Partial.pyx:234-250, Kron.pyx:155-183, plus the
    exact Diag-factor shortcut of Product) against the dense reference matrix."""
    rng = np.random.default_rng(11)
    d = (rng.standard_normal(64) + 1j * rng.standard_normal(64)).astype(np.complex128)
    rows = np.sort(rng.choice(64, 20, replace=False))
    cols = np.sort(rng.choice(64, 33, replace=False))
    mats = [fm.Partial(fm.Fourier(64), cols=cols), fm.Partial(fm.Hadamard(6), rows=rows),
            fm.Partial(fm.Fourier(64), rows=rows, cols=cols), fm.Kron(fm.Fourier(8), fm.Hadamard(3)),
            fm.Product(fm.Partial(fm.Fourier(64), rows=rows), fm.Diag(d)), fm.Product(fm.Diag(d), fm.Fourier(64)),
            fm.Product(fm.Diag(d), fm.Partial(fm.Fourier(64), cols=cols), 2.5)]
    for M in mats:
        ref = M.reference().to(torch.complex128)
        assert ref.shape == (M.numRows, M.numCols)
        cn = torch.linalg.vector_norm(ref, dim=0)
        rn = torch.linalg.vector_norm(ref, dim=1)
        assert float((M.colNorms.to(torch.float64) - cn).abs().max()) <= 1e-9 * float(cn.max()), repr(M)
        assert float((M.rowNorms.to(torch.float64) - rn).abs().max()) <= 1e-9 * float(rn.max()), repr(M)
        s = float(torch.linalg.matrix_norm(ref, ord=2))
        assert abs(M.largestSingularValue - s) <= 1e-8 * s, repr(M)


@pytest.mark.parametrize('tag', ['cs', 'had'])
def test_stela_matches_reference(fm, tag):
    A = build(fm, tag, np.complex128)
    lam = float(GA[tag + '_params'][0])
    b = GA[tag + '_b']
    got = fm.algorithms.STELA(A, numLambda=lam, numMaxSteps=8).process(colmajor(b)).cpu().numpy()
    ref = GA[tag + '_stela']
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-7 * np.abs(ref).max()
