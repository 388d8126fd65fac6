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
    exact Diag-factor shortcut of Product) against dense matrices assembled from the ORACLE's constructions
    (oracle.fastmat_oracle.dense_fourier / dense_hadamard - the reference's own `_reference` recipes), not the package's."""
    from oracle import fastmat_oracle as orc
    rng = np.random.default_rng(11)
    d = (rng.standard_normal(64) + 1j * rng.standard_normal(64)).astype(np.complex128)
    rows = np.sort(rng.choice(64, 20, replace=False))
    cols = np.sort(rng.choice(64, 33, replace=False))
    F64, H6 = orc.dense_fourier(64), orc.dense_hadamard(6).astype(np.float64)
    cases = [(fm.Partial(fm.Fourier(64), cols=cols), F64[:, cols]),
             (fm.Partial(fm.Hadamard(6), rows=rows), H6[rows, :]),
             (fm.Partial(fm.Fourier(64), rows=rows, cols=cols), F64[np.ix_(rows, cols)]),
             (fm.Kron(fm.Fourier(8), fm.Hadamard(3)), np.kron(orc.dense_fourier(8), orc.dense_hadamard(3).astype(np.float64))),
             (fm.Product(fm.Partial(fm.Fourier(64), rows=rows), fm.Diag(d)), F64[rows, :] * d[None, :]),
             (fm.Product(fm.Diag(d), fm.Fourier(64)), d[:, None] * F64),
             (fm.Product(fm.Diag(d), fm.Partial(fm.Fourier(64), cols=cols), 2.5), 2.5 * d[:, None] * F64[:, cols])]
    for M, ref in cases:
        assert ref.shape == (M.numRows, M.numCols)
        # the package's own dense construction agrees with the oracle's, too
        assert np.abs(M.reference().cpu().numpy() - ref).max() <= 1e-9 * np.abs(ref).max(), repr(M)
        cn, rn = np.linalg.norm(ref, axis=0), np.linalg.norm(ref, axis=1)
        assert np.abs(M.colNorms.cpu().numpy().astype(np.float64) - cn).max() <= 1e-9 * cn.max(), repr(M)
        assert np.abs(M.rowNorms.cpu().numpy().astype(np.float64) - rn).max() <= 1e-9 * rn.max(), repr(M)
        s_ref = np.linalg.norm(ref, ord=2)
        assert abs(M.largestSingularValue - s_ref) <= 1e-8 * s_ref, repr(M)


def test_abs_argmax_kernel_matches_numpy(fm):
    """fmb_abs_argmax (OMP's atom selection in one sweep) == np.argmax(np.abs(x), axis=0), first maximum on ties, for every
    dtype, ragged row counts (chunk boundaries) and strided column batches."""
    from fastmat_b200.algorithms.OMP import abs_argmax
    rng = np.random.default_rng(21)
    for dt in (np.float32, np.float64, np.complex64, np.complex128):
        for rows, cols in ((1, 3), (7, 5), (4099, 17), (65536, 9)):
            x = rng.standard_normal((rows, cols))
            if np.dtype(dt).kind == 'c':
                x = x + 1j * rng.standard_normal((rows, cols))
            x = x.astype(dt)
            if rows > 8:                                             # exact ties: the first one must win
                x[5, 0] = x[rows - 3, 0] = 100.0
                x[rows // 2, 1] = x[rows // 2 + 1, 1] = -50.0
            got = abs_argmax(colmajor(x)).cpu().numpy()
            assert np.array_equal(got, np.argmax(np.abs(x), axis=0)), (dt, rows, cols)
    wide = colmajor(rng.standard_normal((300, 12)).astype(np.float32))
    assert np.array_equal(abs_argmax(wide[:, 2:9]).cpu().numpy(), np.argmax(np.abs(wide.cpu().numpy()[:, 2:9]), axis=0))


def test_algorithm_base_parameters_trace_and_callbacks(fm):
    """Behaviour of fastmat/algorithms/Algorithm.pyx the solvers inherit: declared defaults, AttributeError for unknown
    parameters (at construction and at process()), per-step callbacks, snapshot() records in .trace, cbResult once."""
    A = build(fm, 'cs', np.complex128)
    alg = fm.algorithms.ISTA(A)
    assert (alg.numLambda, alg.numMaxSteps, alg.cbStep, alg.cbTrace, alg.cbResult) == (0.1, 100, None, None, None)
    with pytest.raises(AttributeError):
        fm.algorithms.OMP(A, numLamda=3)
    with pytest.raises(NotImplementedError):
        fm.algorithms.Algorithm()
    with pytest.raises(TypeError):
        alg.trace = 'not a list'
    seen, done = [], []
    alg = fm.algorithms.FISTA(A, numMaxSteps=5, cbStep=lambda a: seen.append(a.numStep), cbTrace=fm.algorithms.Algorithm.snapshot,
                              cbResult=lambda a: done.append(a.numStep))
    alg.process(colmajor(GA['cs_b']), numLambda=0.25)
    assert seen == [0, 1, 2, 3, 4] and done == [4] and alg.numLambda == 0.25
    assert [r.numStep for r in alg.trace] == [0, 1, 2, 3, 4] and not hasattr(alg.trace[0], '_trace')
    assert alg.trace[1].arrX is not alg.trace[3].arrX                # records keep the state of THEIR step


@pytest.mark.parametrize('tag', ['cs', 'had'])
def test_stela_matches_reference(fm, tag):
    A = build(fm, tag, np.complex128)
    lam = float(GA[tag + '_params'][0])
    b = GA[tag + '_b']
    got = fm.algorithms.STELA(A, numLambda=lam, numMaxSteps=8).process(colmajor(b)).cpu().numpy()
    ref = GA[tag + '_stela']
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 1e-7 * np.abs(ref).max()
