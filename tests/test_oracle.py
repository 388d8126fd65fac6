"""CPU tests: the numpy oracle (oracle/fastmat_oracle.py) against the fixtures frozen from the REAL reference."""
import numpy as np
import pytest

from conftest import golden_data, relerr, seeded, seeded_typed
from oracle import fastmat_oracle as orc

G = golden_data()
TOL = 2e-12          # double-precision agreement oracle <-> reference (both pocketfft in complex128)


def test_planner_optimal_size_bit_exact():
    p = G.meta['planner']
    for ms in (2, 3, 4, 5, 7):
        got = [orc.find_optimal_fft_size(o, ms) for o in p['orders']]
        assert got == p['opt_%d' % ms]


def test_planner_complexity_bit_exact():
    p = G.meta['planner']
    got = [int(np.float32(orc.get_fft_complexity(o)).view(np.uint32)) for o in p['orders']]
    assert got == p['complexity_bits']


def test_planner_derived_decisions():
    p = G.meta['planner']
    assert [orc.fourier_bluestein_size(o) for o in p['fourier_orders']] == p['fourier_numL']
    assert [orc.circulant_inner_size(n) for n in p['circulant_n']] == p['circulant_inner']
    assert [orc.toeplitz_inner_size(n, m) for n, m in p['toeplitz_nm']] == p['toeplitz_inner']


def test_planner_survey_known_answers():
    # SURVEY.md appendix B
    assert orc.find_optimal_fft_size(2 ** 20 - 1, 4) == 2 ** 20
    assert orc.find_optimal_fft_size(2 * 1000003 - 1, 4) == 2 ** 21
    assert orc.find_optimal_fft_size(2 * 41 - 1, 4) == 96
    assert orc.find_optimal_fft_size(2 ** 24 + 1, 4) == 2 ** 24          # the float32 quirk
    assert float(orc.get_fft_complexity(2 ** 20)) == 53477376.0
    assert float(orc.get_fft_complexity(2 ** 19)) == 25690112.0
    assert orc.fourier_bluestein_size(1000003) == 2 ** 21
    assert orc.fourier_bluestein_size(127) == 256
    assert orc.circulant_inner_size(41) == 96
    assert orc.toeplitz_inner_size(2 ** 19, 2 ** 19) == 2 ** 20


@pytest.mark.parametrize('name', G.cases('fourier'))
def test_fourier(name):
    p = G.params(name)
    x = G.get(name, 'x')
    assert relerr(orc.fourier_forward(x, p['optimize']), G.get(name, 'fwd')) < TOL
    assert relerr(orc.fourier_backward(x, p['optimize']), G.get(name, 'bwd')) < TOL
    if x.ndim == 2 and p['n'] <= 300:
        # dense ground truth, independent of np.fft (the reference's own criterion)
        assert relerr(orc.dense_fourier(p['n']).dot(x), G.get(name, 'fwd')) < 1e-11
        assert relerr(orc.dense_fourier(p['n']).conj().T.dot(x), G.get(name, 'bwd')) < 1e-11


@pytest.mark.parametrize('name', [n for n in G.cases('fourier_big') if G.params(n)['n'] <= 2 ** 18])
def test_fourier_big(name):
    p = G.params(name)
    x = seeded(p['seed'], p['n'], p['cols'])
    assert orc.fourier_bluestein_size(p['n'], p['optimize']) == p['numL']
    for d, f in (('fwd', orc.fourier_forward), ('bwd', orc.fourier_backward)):
        y = f(x, p['optimize'])
        assert relerr(y[G.get(name, d + '_rows')], G.get(name, d)) < TOL
        assert relerr(y.sum(axis=0), G.get(name, d + '_sum')) < 1e-9


@pytest.mark.parametrize('name', G.cases('circulant'))
def test_circulant(name):
    p = G.params(name)
    c, x = G.get(name, 'c'), G.get(name, 'x')
    opt = p.get('optimize', True)
    assert orc.circulant_inner_size(p['n'], opt) == p['inner']
    assert relerr(orc.circulant_forward(c, x, opt), G.get(name, 'fwd')) < TOL
    assert relerr(orc.circulant_backward(c, x, opt), G.get(name, 'bwd')) < TOL
    assert relerr(orc.dense_circulant(c).dot(x), G.get(name, 'fwd')) < 1e-11
    assert relerr(orc.dense_circulant(c).conj().T.dot(x), G.get(name, 'bwd')) < 1e-11


@pytest.mark.parametrize('name', [n for n in G.cases('circulant_big') if G.params(n)['n'] <= 2 ** 17])
def test_circulant_big(name):
    p = G.params(name)
    c = seeded(p['seed_c'], p['n'])
    x = seeded(p['seed'], p['n'], p['cols'])
    assert orc.circulant_inner_size(p['n']) == p['inner']
    for d, f in (('fwd', orc.circulant_forward), ('bwd', orc.circulant_backward)):
        y = f(c, x)
        assert relerr(y[G.get(name, d + '_rows')], G.get(name, d)) < TOL


@pytest.mark.parametrize('name', G.cases('toeplitz'))
def test_toeplitz(name):
    p = G.params(name)
    vc, vr, x, y = (G.get(name, k) for k in ('vc', 'vr', 'x', 'y'))
    assert orc.toeplitz_inner_size(p['n'], p['m']) == p['inner']
    assert relerr(orc.toeplitz_forward(vc, vr, x), G.get(name, 'fwd')) < TOL
    assert relerr(orc.toeplitz_backward(vc, vr, y), G.get(name, 'bwd')) < TOL
    assert relerr(orc.dense_toeplitz(vc, vr).dot(x), G.get(name, 'fwd')) < 1e-11
    assert relerr(orc.dense_toeplitz(vc, vr).conj().T.dot(y), G.get(name, 'bwd')) < 1e-11


@pytest.mark.parametrize('name', [n for n in G.cases('toeplitz_big') if G.params(n)['n'] <= 2 ** 17])
def test_toeplitz_big(name):
    p = G.params(name)
    vc = seeded(p['seed_c'], p['n'])
    vr = seeded(p['seed_r'], p['m'] - 1)
    x = seeded(p['seed_x'], p['m'], p['cols'])
    y = seeded(p['seed_y'], p['n'], p['cols'])
    assert orc.toeplitz_inner_size(p['n'], p['m']) == p['inner']
    f = orc.toeplitz_forward(vc, vr, x)
    b = orc.toeplitz_backward(vc, vr, y)
    assert relerr(f[G.get(name, 'fwd_rows')], G.get(name, 'fwd')) < TOL
    assert relerr(b[G.get(name, 'bwd_rows')], G.get(name, 'bwd')) < TOL


@pytest.mark.parametrize('name', G.cases('hadamard'))
def test_hadamard_bit_exact(name):
    x = G.get(name, 'x')
    y = orc.hadamard_forward(x)
    ref = G.get(name, 'fwd')
    assert y.dtype == ref.dtype
    assert np.array_equal(y.view(np.uint8), np.ascontiguousarray(ref).view(np.uint8))     # bit-exact, all 8 dtypes


@pytest.mark.parametrize('name', [n for n in G.cases('hadamard_big') if G.params(n)['order'] <= 16])
def test_hadamard_big_bit_exact(name):
    p = G.params(name)
    x = seeded_typed(p['seed'], p['dtype'], 2 ** p['order'], p['cols'])
    y = orc.hadamard_forward(x)
    ref = G.get(name, 'fwd')
    got = np.ascontiguousarray(y[G.get(name, 'fwd_rows')])
    assert got.dtype == ref.dtype
    assert np.array_equal(got.view(np.uint8), np.ascontiguousarray(ref).view(np.uint8))


def test_hadamard_dense_small():
    for order in (1, 3, 5):
        x = np.random.default_rng(order).integers(-5, 5, size=(2 ** order, 3)).astype(np.int32)
        assert np.array_equal(orc.hadamard_forward(x), orc.dense_hadamard(order, np.int32).dot(x))


@pytest.mark.parametrize('name', G.cases('kron_fourier'))
def test_kron_fourier(name):
    p = G.params(name)
    x = G.get(name, 'x')
    assert relerr(orc.kron_fourier_forward(p['dims'], x), G.get(name, 'fwd')) < TOL
    assert relerr(orc.kron_fourier_backward(p['dims'], x), G.get(name, 'bwd')) < TOL
    # and the literal reshape choreography of Kron.pyx:267-303
    applies = [lambda a: np.fft.fft(a, axis=0)] * len(p['dims'])
    assert relerr(orc.kron_forward(applies, p['dims'], x.astype(np.complex128)), G.get(name, 'fwd')) < TOL


def test_kron_dense_and_mixed():
    name = 'kron_dense_5x4x3'
    a = [G.get(name, 'a%d' % i) for i in range(3)]
    x = G.get(name, 'x')
    y = orc.kron_forward([lambda v, m=m: m.dot(v) for m in a], [5, 4, 3], x)
    assert relerr(y, G.get(name, 'fwd')) < TOL
    assert relerr(np.kron(np.kron(a[0], a[1]), a[2]).dot(x), G.get(name, 'fwd')) < TOL
    name = 'kron_H3_F5'
    x = G.get(name, 'x')
    y = orc.kron_forward([lambda v: orc.hadamard_forward(v), lambda v: orc.fourier_forward(v)], [8, 5],
                         x.astype(np.complex128))
    assert relerr(y, G.get(name, 'fwd')) < TOL


def test_partial_hadamard():
    name = 'partial_hadamard4'
    rows, cols, x, y = (G.get(name, k) for k in ('rows', 'cols', 'x', 'y'))
    f = orc.partial_forward(orc.hadamard_forward, 16, rows, cols, x)
    b = orc.partial_backward(orc.hadamard_forward, 16, rows, cols, y)
    assert np.array_equal(f, G.get(name, 'fwd')) and f.dtype == G.get(name, 'fwd').dtype
    assert np.array_equal(b, G.get(name, 'bwd'))
    name = 'partial_hadamard4_bool'
    rows = np.arange(16)[G.get(name, 'rows')]
    f = orc.partial_forward(orc.hadamard_forward, 16, rows, None, G.get(name, 'x'))
    b = orc.partial_backward(orc.hadamard_forward, 16, rows, None, G.get(name, 'y'))
    assert np.array_equal(f, G.get(name, 'fwd'))
    assert np.array_equal(b, G.get(name, 'bwd'))


def test_cs_operator():
    name = 'cs_partial_fourier_diag'
    idx, d, x, y = (G.get(name, k) for k in ('idx', 'd', 'x', 'y'))
    f = orc.partial_forward(orc.fourier_forward, 256, idx, None, orc.diag_forward(d, x))
    b = orc.diag_backward(d, orc.partial_backward(orc.fourier_backward, 256, idx, None, y))
    assert relerr(f, G.get(name, 'fwd')) < TOL
    assert relerr(b, G.get(name, 'bwd')) < TOL


@pytest.mark.parametrize('name', G.cases('diag'))
def test_diag(name):
    d, x = G.get(name, 'd'), G.get(name, 'x')
    f, b = orc.diag_forward(d, x), orc.diag_backward(d, x)
    rf, rb = G.get(name, 'fwd'), G.get(name, 'bwd')
    assert f.dtype == rf.dtype and b.dtype == rb.dtype
    if np.issubdtype(rf.dtype, np.integer):
        assert np.array_equal(f, rf) and np.array_equal(b, rb)
    else:
        assert relerr(f, rf) < 1e-6 and relerr(b, rb) < 1e-6


def test_permutation_bit_exact():
    name = 'permutation_35'
    sigma, x = G.get(name, 'sigma'), G.get(name, 'x')
    assert np.array_equal(orc.permutation_forward(sigma, x), G.get(name, 'fwd'))
    assert np.array_equal(orc.permutation_backward(sigma, x), G.get(name, 'bwd'))


def test_reference_dtype_record():
    # as-is behaviour of the reference on this numpy (SURVEY.md key fact 6) - recorded, and relied on by the
    # dtype policy documented in DESIGN.md
    d = G.meta['dtypes']
    assert d['Fourier']['complex64'] == 'complex64' and d['Fourier']['float64'] == 'complex128'
    assert all(d['Hadamard'][t] == t for t in d['Hadamard'])
    assert d['Circulant_c64']['complex64'] == 'complex128'
