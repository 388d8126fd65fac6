"""GPU parity tests: the CUDA path, called through the class layer -> C-ABI (libfastmat_b200.so), against
(a) the fixtures frozen from the real reference (tests/golden) and (b) the numpy oracle on seeded inputs, plus
size-independent properties at the BASELINE sizes.

Tolerance (BASELINE.json north_star): max|y - y_ref| <= tol * ||x||_2 * log2(N) per column with tol = 1e-5 for
complex64/float32 and 1e-12 for complex128/float64; integer Hadamard / Permutation / Partial are bit-exact.
"""
import os

import numpy as np
import pytest
import torch

from conftest import golden_data, seeded, seeded_typed

pytestmark = pytest.mark.gpu

G = golden_data()
TOL64, TOL128 = 1e-5, 1e-12


@pytest.fixture(scope='module')
def fm():
    import fastmat_b200
    assert torch.cuda.is_available()
    return fastmat_b200


@pytest.fixture(scope='module')
def orc():
    from oracle import fastmat_oracle
    return fastmat_oracle


def dev(a, layout='F'):
    """numpy (n, M) / (n,) -> CUDA tensor in column-major ('F', fastmat native) or row-major ('C') layout."""
    a = np.asarray(a)
    if a.ndim == 1:
        return torch.from_numpy(a.copy()).cuda()
    if layout == 'C':
        return torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()


def host(t):
    return t.detach().cpu().numpy()


def nerr(y, ref, x, n):
    """Error normalised by ||x|| log2 N (per column, worst column)."""
    y, ref, x = np.asarray(y), np.asarray(ref), np.asarray(x)
    if y.ndim == 1:
        y, ref, x = y[:, None], ref[:, None], x[:, None]
    nx = np.linalg.norm(x.astype(np.complex128), axis=0)
    nx[nx == 0] = 1.0
    return float((np.abs(y - ref).max(axis=0) / (nx * max(1.0, np.log2(max(n, 2))))).max())


def check_pair(M, x, ref_f, y, ref_b, n, both_precisions=True, layouts=('F', 'C')):
    """forward(x) ~ ref_f and backward(y) ~ ref_b in complex128 and (cast) complex64, both layouts."""
    for lay in layouts:
        f = M.forward(dev(x, lay))
        b = M.backward(dev(y, lay))
        assert f.dtype == torch.complex128 and b.dtype == torch.complex128
        assert nerr(host(f), ref_f, x, n) < TOL128, ('fwd c128', lay)
        assert nerr(host(b), ref_b, y, n) < TOL128, ('bwd c128', lay)
        if both_precisions:
            x32, y32 = x.astype(np.complex64), y.astype(np.complex64)
            f = M.forward(dev(x32, lay))
            b = M.backward(dev(y32, lay))
            assert nerr(host(f), ref_f, x, n) < TOL64, ('fwd c64', lay)
            assert nerr(host(b), ref_b, y, n) < TOL64, ('bwd c64', lay)


# ------------------------------------------------------------------------------------------- Fourier
@pytest.mark.parametrize('name', G.cases('fourier'))
def test_fourier_golden(fm, name):
    p = G.params(name)
    x = G.get(name, 'x')
    F = fm.Fourier(p['n'], optimize=p['optimize'])
    check_pair(F, x, G.get(name, 'fwd'), x, G.get(name, 'bwd'), p['n'], layouts=('F', 'C') if x.ndim == 2 else ('F', ))
    out = F.forward(dev(x.astype(np.complex64)))
    assert out.dtype == torch.complex64 and out.ndim == x.ndim


@pytest.mark.parametrize('name', G.cases('fourier_big'))
def test_fourier_big_golden(fm, name):
    p = G.params(name)
    n = p['n']
    x = seeded(p['seed'], n, p['cols'])
    F = fm.Fourier(n, optimize=p['optimize'])
    assert F._numL == p['numL']                       # the reference's Bluestein decision is reproduced
    # When the reference takes its chirp-z branch it forms k^2 * pi / N in float64 (fastmat/Fourier.pyx:130), which
    # costs it ~1e-10 at N = 1e6; the device path reduces k^2 mod 2N in integers first and is MORE accurate.  Those
    # cases are therefore compared with the reference at 1e-9 and, at full tolerance, with pocketfft's direct DFT.
    direct = None
    if p['numL'] > 0:
        direct = {'fwd': np.fft.fft(x, axis=0), 'bwd': np.conj(np.fft.fft(np.conj(x), axis=0))}
    for lay in ('F', 'C'):
        for dt, tol in ((np.complex128, TOL128), (np.complex64, TOL64)):
            xd = dev(x.astype(dt), lay)
            keep = xd.clone()
            for d, y in (('fwd', F.forward(xd)), ('bwd', F.backward(xd))):
                rows = G.get(name, d + '_rows')
                yh = host(y)
                if direct is not None:
                    assert nerr(yh, direct[d], x, n) < tol, (d, lay, dt, 'vs direct DFT')
                    assert nerr(yh[rows], G.get(name, d), x, n) < max(tol, 1e-9), (d, lay, dt)
                else:
                    assert nerr(yh[rows], G.get(name, d), x, n) < tol, (d, lay, dt)
                s = np.abs(yh.sum(axis=0) - G.get(name, d + '_sum')).max()
                assert s / (np.linalg.norm(x, axis=0).max() * np.sqrt(n) * np.log2(n)) < tol * 10
            assert torch.equal(xd, keep)              # the input is never modified (inspect/test.py:334-338)


def test_fourier_dense_small(fm):
    # the reference's own criterion: compare against the dense DFT matrix (no FFT involved), N in {35, 127}
    for n in (35, 127, 64, 243):
        F = fm.Fourier(n)
        x = seeded(n, n, 5)
        ref = host(F.reference()).dot(x)
        assert nerr(host(F.forward(dev(x))), ref, x, n) < TOL128
        assert nerr(host(F * dev(x)), ref, x, n) < TOL128             # operator interface


def test_fourier_real_and_int_inputs(fm):
    n = 256
    F = fm.Fourier(n)
    rng = np.random.default_rng(3)
    for dt, out_dt, tol in ((np.float32, torch.complex64, TOL64), (np.float64, torch.complex128, TOL128),
                            (np.int8, torch.complex64, TOL64), (np.int16, torch.complex64, TOL64),
                            (np.int32, torch.complex128, TOL128), (np.int64, torch.complex128, TOL128)):
        x = (rng.standard_normal((n, 3)) * 50).astype(dt)
        y = F.forward(dev(x))
        assert y.dtype == out_dt
        assert nerr(host(y), np.fft.fft(x.astype(np.float64), axis=0), x, n) < tol


def test_fourier_roundtrip_full_size(fm):
    # property at the BASELINE size: backward(forward(x)) = N x; Parseval
    n, m = 2 ** 20, 8
    F = fm.Fourier(n)
    g = torch.Generator(device='cuda').manual_seed(1234)
    x = torch.randn((m, n), dtype=torch.float32, device='cuda', generator=g).to(torch.complex64).t()
    y = F.forward(x)
    z = F.backward(y)
    err = (z / n - x).abs().max().item() / x.abs().max().item()
    assert err < 1e-4
    e_in = torch.linalg.vector_norm(x, dim=0) ** 2
    e_out = torch.linalg.vector_norm(y, dim=0) ** 2 / n
    assert float(((e_in - e_out).abs() / e_in).max()) < 1e-4


# ------------------------------------------------------------------------------------------- Circulant
@pytest.mark.parametrize('name', G.cases('circulant'))
def test_circulant_golden(fm, name):
    p = G.params(name)
    c, x = G.get(name, 'c'), G.get(name, 'x')
    C = fm.Circulant(c, optimize=p.get('optimize', True))
    check_pair(C, x, G.get(name, 'fwd'), x, G.get(name, 'bwd'), max(p['n'], 2))


@pytest.mark.parametrize('name', G.cases('circulant_big'))
def test_circulant_big_golden(fm, name):
    p = G.params(name)
    n = p['n']
    c = seeded(p['seed_c'], n)
    x = seeded(p['seed'], n, p['cols'])
    for cdt, xdt, tol in ((np.complex128, np.complex128, TOL128), (np.complex64, np.complex64, TOL64)):
        C = fm.Circulant(c.astype(cdt))
        for lay in ('F', 'C'):
            xd = dev(x.astype(xdt), lay)
            for d, y in (('fwd', C.forward(xd)), ('bwd', C.backward(xd))):
                assert y.dtype == (torch.complex128 if xdt == np.complex128 else torch.complex64)
                rows = G.get(name, d + '_rows')
                # ||C x|| scales with ||c||: normalise by ||c||_2 as well
                e = nerr(host(y)[rows], G.get(name, d), x, n) / np.linalg.norm(c)
                assert e < tol, (d, lay, e)


@pytest.mark.parametrize('name', G.cases('circulant_ml'))
def test_circulant_multilevel_golden(fm, name):
    c, x = G.get(name, 'c'), G.get(name, 'x')
    C = fm.Circulant(c)
    n = x.shape[0]
    assert nerr(host(C.forward(dev(x))), G.get(name, 'fwd'), x, n) / np.linalg.norm(c) < TOL128
    assert nerr(host(C.backward(dev(x))), G.get(name, 'bwd'), x, n) / np.linalg.norm(c) < TOL128
    ref = host(C.reference())
    assert nerr(ref.dot(x), G.get(name, 'fwd'), x, n) / np.linalg.norm(c) < TOL128


def test_circulant_properties_full_size(fm, orc):
    # BASELINE config 2 shape (N = 2^20, complex64): adjoint identity, linearity, and two columns against the oracle
    n, m = 2 ** 20, 16
    rng = np.random.default_rng(4321)
    c = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    C = fm.Circulant(c)
    g = torch.Generator(device='cuda').manual_seed(1234)
    x = torch.complex(torch.randn((m, n), device='cuda', generator=g), torch.randn((m, n), device='cuda', generator=g)).t()
    y = torch.complex(torch.randn((m, n), device='cuda', generator=g), torch.randn((m, n), device='cuda', generator=g)).t()
    cx = C.forward(x)
    chy = C.backward(y)
    lhs = (cx.conj() * y).sum(dim=0)                  # <C x, y>
    rhs = (x.conj() * chy).sum(dim=0)                 # <x, C^H y>
    scale = torch.linalg.vector_norm(cx, dim=0) * torch.linalg.vector_norm(y, dim=0)
    assert float(((lhs - rhs).abs() / scale).max()) < 1e-4
    lin = C.forward(2 * x + y) - (2 * cx + C.forward(y))
    assert float(lin.abs().max() / cx.abs().max()) < 1e-4
    xh = host(x[:, :2])
    ref = orc.circulant_forward(c, xh)
    assert nerr(host(cx[:, :2]), ref, xh, n) / np.linalg.norm(c.astype(np.complex128)) < TOL64
    refb = orc.circulant_backward(c, xh)
    assert nerr(host(C.backward(x[:, :2])), refb, xh, n) / np.linalg.norm(c.astype(np.complex128)) < TOL64


# ------------------------------------------------------------------------------------------- Toeplitz
@pytest.mark.parametrize('name', G.cases('toeplitz'))
def test_toeplitz_golden(fm, name):
    p = G.params(name)
    vc, vr, x, y = (G.get(name, k) for k in ('vc', 'vr', 'x', 'y'))
    T = fm.Toeplitz(vc, vr)
    assert T.shape == (p['n'], p['m'])
    check_pair(T, x, G.get(name, 'fwd'), y, G.get(name, 'bwd'), max(p['n'] + p['m'], 2))
    assert np.array_equal(host(T.reference()), __import__('oracle.fastmat_oracle', fromlist=['x']).dense_toeplitz(vc, vr))


@pytest.mark.parametrize('name', G.cases('toeplitz_big'))
def test_toeplitz_big_golden(fm, name):
    p = G.params(name)
    n, m = p['n'], p['m']
    vc = seeded(p['seed_c'], n)
    vr = seeded(p['seed_r'], m - 1)
    x = seeded(p['seed_x'], m, p['cols'])
    y = seeded(p['seed_y'], n, p['cols'])
    nt = np.sqrt(np.linalg.norm(vc) ** 2 + np.linalg.norm(vr) ** 2)
    for dt, tol in ((np.complex128, TOL128), (np.complex64, TOL64)):
        T = fm.Toeplitz(vc.astype(dt), vr.astype(dt))
        for lay in ('F', 'C'):
            f = T.forward(dev(x.astype(dt), lay))
            b = T.backward(dev(y.astype(dt), lay))
            assert f.shape == (n, p['cols']) and b.shape == (m, p['cols'])
            assert nerr(host(f)[G.get(name, 'fwd_rows')], G.get(name, 'fwd'), x, n + m) / nt < tol
            assert nerr(host(b)[G.get(name, 'bwd_rows')], G.get(name, 'bwd'), y, n + m) / nt < tol


@pytest.mark.parametrize('name', G.cases('toeplitz_ml'))
def test_toeplitz_multilevel_golden(fm, name):
    p = G.params(name)
    t, x, y = G.get(name, 't'), G.get(name, 'x'), G.get(name, 'y')
    T = fm.Toeplitz(t, split=p['split']) if p['split'] is not None else fm.Toeplitz(t)
    n = t.size
    assert nerr(host(T.forward(dev(x))), G.get(name, 'fwd'), x, n) / np.linalg.norm(t) < TOL128
    assert nerr(host(T.backward(dev(y))), G.get(name, 'bwd'), y, n) / np.linalg.norm(t) < TOL128
    assert nerr(host(T.reference()).dot(x), G.get(name, 'fwd'), x, n) / np.linalg.norm(t) < TOL128


def test_toeplitz_full_size_adjoint(fm):
    n = m = 2 ** 19
    rng = np.random.default_rng(4321)
    vc = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    vr = (rng.standard_normal(m - 1) + 1j * rng.standard_normal(m - 1)).astype(np.complex64)
    T = fm.Toeplitz(vc, vr)
    assert T._plan.info.inner_size == 2 ** 20
    g = torch.Generator(device='cuda').manual_seed(99)
    x = torch.complex(torch.randn((8, m), device='cuda', generator=g), torch.randn((8, m), device='cuda', generator=g)).t()
    y = torch.complex(torch.randn((8, n), device='cuda', generator=g), torch.randn((8, n), device='cuda', generator=g)).t()
    tx, thy = T.forward(x), T.backward(y)
    lhs = (tx.conj() * y).sum(dim=0)
    rhs = (x.conj() * thy).sum(dim=0)
    scale = torch.linalg.vector_norm(tx, dim=0) * torch.linalg.vector_norm(y, dim=0)
    assert float(((lhs - rhs).abs() / scale).max()) < 1e-4


# ------------------------------------------------------------------------------------------- Hadamard
@pytest.mark.parametrize('name', G.cases('hadamard'))
def test_hadamard_golden_bit_exact(fm, name):
    p = G.params(name)
    x = G.get(name, 'x')
    ref = G.get(name, 'fwd')
    H = fm.Hadamard(p['order'])
    for lay in ('F', 'C'):
        xd = dev(x, lay)
        keep = xd.clone()
        y = host(H.forward(xd))
        assert y.dtype == ref.dtype
        assert np.array_equal(np.ascontiguousarray(y).view(np.uint8), np.ascontiguousarray(ref).view(np.uint8)), lay
        yb = host(H.backward(xd))
        assert np.array_equal(np.ascontiguousarray(yb).view(np.uint8), np.ascontiguousarray(ref).view(np.uint8))
        assert torch.equal(xd, keep)


@pytest.mark.parametrize('name', G.cases('hadamard_big'))
def test_hadamard_big_golden_bit_exact(fm, name):
    p = G.params(name)
    x = seeded_typed(p['seed'], p['dtype'], 2 ** p['order'], p['cols'])
    H = fm.Hadamard(p['order'])
    ref = G.get(name, 'fwd')
    for lay in ('F', 'C'):
        y = host(H.forward(dev(x, lay)))
        got = np.ascontiguousarray(y[G.get(name, 'fwd_rows')])
        assert np.array_equal(got.view(np.uint8), np.ascontiguousarray(ref).view(np.uint8)), lay
        if np.issubdtype(x.dtype, np.integer):
            with np.errstate(over='ignore'):
                assert np.array_equal(y.sum(axis=0, dtype=x.dtype), G.get(name, 'fwd_sum').astype(x.dtype))


def test_hadamard_involution_full_size(fm):
    # BASELINE config 3 shape: H(H(x)) = N x exactly for int32 (wrap-around arithmetic included)
    order, m = 20, 8
    H = fm.Hadamard(order)
    g = torch.Generator(device='cuda').manual_seed(5)
    x = torch.randint(-2 ** 31, 2 ** 31 - 1, (m, 2 ** order), dtype=torch.int32, device='cuda', generator=g).t()
    z = H.forward(H.forward(x))
    assert torch.equal(z, x * (2 ** order))           # int32 arithmetic wraps identically on both sides
    xf = torch.randn((m, 2 ** order), dtype=torch.float32, device='cuda', generator=g).t()
    zf = H.forward(H.forward(xf))
    assert float((zf / 2 ** order - xf).abs().max()) < 1e-3


# ------------------------------------------------------------------------------------------- Kron
@pytest.mark.parametrize('name', G.cases('kron_fourier'))
def test_kron_fourier_golden(fm, name):
    p = G.params(name)
    x = G.get(name, 'x')
    K = fm.Kron(*[fm.Fourier(d) for d in p['dims']])
    assert (K._plan is not None) == (len(p['dims']) == 2)
    check_pair(K, x, G.get(name, 'fwd'), x, G.get(name, 'bwd'), x.shape[0])


def test_kron_fourier_big_golden(fm):
    name = 'kronF_big_1024x1024'
    p = G.params(name)
    x = seeded(p['seed'], 2 ** 20, p['cols'])
    K = fm.Kron(fm.Fourier(1024), fm.Fourier(1024))
    for dt, tol in ((np.complex128, TOL128), (np.complex64, TOL64)):
        for d, y in (('fwd', K.forward(dev(x.astype(dt)))), ('bwd', K.backward(dev(x.astype(dt))))):
            assert nerr(host(y)[G.get(name, d + '_rows')], G.get(name, d), x, 2 ** 20) < tol


def test_kron_mixed_and_dense_golden(fm):
    name = 'kron_H3_F5'
    x = G.get(name, 'x')
    K = fm.Kron(fm.Hadamard(3), fm.Fourier(5))
    assert nerr(host(K.forward(dev(x))), G.get(name, 'fwd'), x, 40) < TOL128
    assert nerr(host(K.backward(dev(x))), G.get(name, 'bwd'), x, 40) < TOL128
    name = 'kron_dense_5x4x3'
    a = [G.get(name, 'a%d' % i) for i in range(3)]
    x = G.get(name, 'x')
    K = fm.Kron(*[fm.Matrix(m) for m in a])
    assert nerr(host(K.forward(dev(x))), G.get(name, 'fwd'), x, 60) < 1e-12 * 60
    assert nerr(host(K.backward(dev(x))), G.get(name, 'bwd'), x, 60) < 1e-12 * 60
    with pytest.raises(ValueError):
        fm.Kron(fm.Fourier(4))
    with pytest.raises(ValueError):
        fm.Kron(fm.Fourier(4), fm.Partial(fm.Fourier(4), rows=np.arange(2)))


# ------------------------------------------------------------------------------------------- glue
def test_partial_hadamard_golden_exact(fm):
    name = 'partial_hadamard4'
    rows, cols, x, y = (G.get(name, k) for k in ('rows', 'cols', 'x', 'y'))
    P = fm.Partial(fm.Hadamard(4), rows=rows, cols=cols)
    assert P.shape == (5, 4)
    assert np.array_equal(host(P.forward(dev(x))), G.get(name, 'fwd'))
    assert np.array_equal(host(P.backward(dev(y))), G.get(name, 'bwd'))
    name = 'partial_hadamard4_bool'
    P = fm.Partial(fm.Hadamard(4), rows=G.get(name, 'rows'))
    assert np.array_equal(host(P.forward(dev(G.get(name, 'x')))), G.get(name, 'fwd'))
    assert np.array_equal(host(P.backward(dev(G.get(name, 'y')))), G.get(name, 'bwd'))
    with pytest.raises(ValueError):
        fm.Partial(fm.Hadamard(4), rows=np.array([1, 16]))
    with pytest.raises(TypeError):
        fm.Partial(fm.Hadamard(4), rows=np.array([0.5, 1.0]))


def test_cs_operator_golden(fm):
    # BASELINE config 5's operator: Product(Partial(Fourier(n), rows=idx), Diag(d))
    name = 'cs_partial_fourier_diag'
    idx, d, x, y = (G.get(name, k) for k in ('idx', 'd', 'x', 'y'))
    A = fm.Product(fm.Partial(fm.Fourier(256), rows=idx), fm.Diag(d))
    assert A.shape == (64, 256)
    assert nerr(host(A.forward(dev(x))), G.get(name, 'fwd'), x, 256) < TOL128
    assert nerr(host(A.backward(dev(y))), G.get(name, 'bwd'), y, 256) < TOL128
    A32 = fm.Product(fm.Partial(fm.Fourier(256), rows=idx), fm.Diag(d.astype(np.complex64)))
    out = A32.forward(dev(x.astype(np.complex64)))
    assert out.dtype == torch.complex64
    assert nerr(host(out), G.get(name, 'fwd'), x, 256) < TOL64


@pytest.mark.parametrize('name', G.cases('diag'))
def test_diag_golden(fm, name):
    d, x = G.get(name, 'd'), G.get(name, 'x')
    D = fm.Diag(d)
    for lay in ('F', 'C'):
        f, b = host(D.forward(dev(x, lay))), host(D.backward(dev(x, lay)))
        rf, rb = G.get(name, 'fwd'), G.get(name, 'bwd')
        assert f.dtype == rf.dtype and b.dtype == rb.dtype
        if np.issubdtype(rf.dtype, np.integer):
            assert np.array_equal(f, rf) and np.array_equal(b, rb)
        else:
            tol = 1e-6 if rf.dtype in (np.float32, np.complex64) else 1e-14
            assert np.abs(f - rf).max() <= tol * np.abs(rf).max()
            assert np.abs(b - rb).max() <= tol * np.abs(rb).max()


def test_permutation_golden_bit_exact(fm):
    name = 'permutation_35'
    sigma, x = G.get(name, 'sigma'), G.get(name, 'x')
    P = fm.Permutation(sigma)
    assert np.array_equal(host(P.forward(dev(x))), G.get(name, 'fwd'))
    assert np.array_equal(host(P.backward(dev(x))), G.get(name, 'bwd'))
    assert np.array_equal(host(P.forward(dev(x, 'C'))), G.get(name, 'fwd'))
    with pytest.raises(ValueError):
        fm.Permutation(np.array([0, 0, 1]))


def test_sum_blocks_product_views_golden(fm):
    name = 'sum_circ_diag_fourier'
    c, d, x = G.get(name, 'c'), G.get(name, 'd'), G.get(name, 'x')
    S = fm.Sum(fm.Circulant(c), fm.Diag(d), fm.Fourier(16))
    assert nerr(host(S.forward(dev(x))), G.get(name, 'fwd'), x, 16) < 1e-12 * 10
    assert nerr(host(S.backward(dev(x))), G.get(name, 'bwd'), x, 16) < 1e-12 * 10
    S2 = fm.Circulant(c) + fm.Diag(d) + fm.Fourier(16)
    assert nerr(host(S2.forward(dev(x))), G.get(name, 'fwd'), x, 16) < 1e-12 * 10
    name = 'blocks_2x2'
    c, d, x = G.get(name, 'c'), G.get(name, 'd'), G.get(name, 'x')
    B = fm.Blocks([[fm.Circulant(c), fm.Fourier(16)], [fm.Diag(d), fm.Hadamard(4)]])
    assert nerr(host(B.forward(dev(x))), G.get(name, 'fwd'), x, 32) < 1e-12 * 10
    assert nerr(host(B.backward(dev(x))), G.get(name, 'bwd'), x, 32) < 1e-12 * 10
    name = 'blockdiag'
    x = G.get(name, 'x')
    BD = fm.BlockDiag(fm.Fourier(16), fm.Hadamard(4))
    assert nerr(host(BD.forward(dev(x))), G.get(name, 'fwd'), x, 32) < 1e-12 * 10
    assert nerr(host(BD.backward(dev(x))), G.get(name, 'bwd'), x, 32) < 1e-12 * 10
    name = 'product_scalar'
    d, x = G.get(name, 'd'), G.get(name, 'x')
    Pr = fm.Product(fm.Hadamard(4), 2.5 - 1j, fm.Diag(d), fm.Fourier(16))
    assert nerr(host(Pr.forward(dev(x))), G.get(name, 'fwd'), x, 16) < 1e-12 * 100
    assert nerr(host(Pr.backward(dev(x))), G.get(name, 'bwd'), x, 16) < 1e-12 * 100
    Pr2 = fm.Hadamard(4) * (2.5 - 1j) * fm.Diag(d) * fm.Fourier(16)
    assert nerr(host(Pr2.forward(dev(x))), G.get(name, 'fwd'), x, 16) < 1e-12 * 100
    name = 'fourier_views'
    x = G.get(name, 'x')
    F = fm.Fourier(16)
    assert nerr(host(F.H.forward(dev(x))), G.get(name, 'H'), x, 16) < TOL128
    assert nerr(host(F.T.forward(dev(x))), G.get(name, 'T'), x, 16) < TOL128
    assert nerr(host(F.conj.forward(dev(x))), G.get(name, 'conj'), x, 16) < TOL128
    assert F.H.H is F and F.conj.conj is F


# ------------------------------------------------------------------------------------------- interface behaviour
def test_dimension_and_type_errors(fm):
    F = fm.Fourier(16)
    with pytest.raises(ValueError):
        F.forward(torch.zeros(15, dtype=torch.complex64, device='cuda'))
    with pytest.raises(ValueError):
        F.forward(torch.zeros((16, 2, 2), dtype=torch.complex64, device='cuda'))
    with pytest.raises(ValueError):
        F.backward(torch.zeros((17, 2), dtype=torch.complex64, device='cuda'))
    with pytest.raises(TypeError):
        F.forward(torch.zeros(16, dtype=torch.float16, device='cuda'))
    with pytest.raises(TypeError):
        F.forward([0.0] * 16)
    with pytest.raises(RuntimeError):
        F.forward(torch.zeros(16, dtype=torch.complex64))                 # CPU tensor: no CPU path
    for bad in (0, -3):
        with pytest.raises(ValueError):
            fm.Fourier(bad)
    for bad in (0, 63):
        with pytest.raises(ValueError):
            fm.Hadamard(bad)
    with pytest.raises(ValueError):
        fm.Toeplitz(np.ones(3), np.ones(2), split=[2])
    with pytest.raises(ValueError):
        fm.Diag(np.ones((2, 2)))


def test_strided_views_and_numpy_convenience(fm, orc):
    n = 1024
    rng = np.random.default_rng(11)
    big = (rng.standard_normal((n, 12)) + 1j * rng.standard_normal((n, 12))).astype(np.complex64)
    t = torch.from_numpy(big).cuda()
    view = t[:, ::3]                                   # step-3 strided columns (inspect/common.py:478-521 "strided")
    F = fm.Fourier(n)
    ref = np.fft.fft(big[:, ::3].astype(np.complex128), axis=0)
    assert nerr(host(F.forward(view)), ref, big[:, ::3], n) < TOL64
    H = fm.Hadamard(10)
    xi = rng.integers(-100, 100, size=(n, 12)).astype(np.int16)
    vi = torch.from_numpy(xi).cuda()[:, 1::4]
    assert np.array_equal(host(H.forward(vi)), orc.hadamard_forward(xi[:, 1::4]))
    # numpy in -> numpy out (host round trip through the same C-ABI call)
    y = F.forward(big[:, 0].copy())
    assert isinstance(y, np.ndarray) and y.shape == (n, )
    assert nerr(y, np.fft.fft(big[:, 0].astype(np.complex128)), big[:, 0], n) < TOL64


def test_zero_columns_and_single_column(fm):
    F = fm.Fourier(64)
    y = F.forward(torch.zeros((64, 0), dtype=torch.complex64, device='cuda'))
    assert y.shape == (64, 0)
    x = torch.ones(64, dtype=torch.complex64, device='cuda')
    y = F.forward(x)
    assert y.shape == (64, ) and abs(y[0].item() - 64) < 1e-4 and float(y[1:].abs().max()) < 1e-4


def test_launches_counted(fm):
    before = fm.launch_count()
    fm.Hadamard(8).forward(torch.ones((256, 4), dtype=torch.float32, device='cuda'))
    assert fm.launch_count() > before


# ------------------------------------------------------------------------------------------- opt-in fused kernels
def _run_with_env(env, code):
    import os
    import subprocess
    import sys
    from conftest import ROOT
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, '-c', code], cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


_SWITCH_CHECK = r'''
import sys, numpy as np, torch
sys.path.insert(0, '.')
import fastmat_b200 as fm
from oracle import fastmat_oracle as orc
rng = np.random.default_rng(5)
def crand(*s): return (rng.standard_normal(s) + 1j * rng.standard_normal(s))
def dev(a, rm=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if rm else t.t().contiguous().t()
n, m = 2 ** 20, 13                                     # enough columns for the pipelined-slab schedule, ragged last slab
x = crand(n, m).astype(np.complex64)
c = crand(n).astype(np.complex64)
C = fm.Circulant(c)
nc = np.linalg.norm(c) * np.linalg.norm(x, axis=0).max() * 20
chk = [0, 7, 12]
for rm in (False, True):
    yf, yb = C.forward(dev(x, rm)).cpu().numpy(), C.backward(dev(x, rm)).cpu().numpy()
    assert np.abs(yf[:, chk] - orc.circulant_forward(c, x[:, chk])).max() / nc < 1e-5
    assert np.abs(yb[:, chk] - orc.circulant_backward(c, x[:, chk])).max() / nc < 1e-5
nt = n // 2
vc, vr = c[:nt].copy(), c[nt:2 * nt - 1].copy()
T = fm.Toeplitz(vc, vr)
xt = x[:nt]
yf, yb = T.forward(dev(xt)).cpu().numpy(), T.backward(dev(xt)).cpu().numpy()
assert np.abs(yf[:, chk] - orc.toeplitz_forward(vc, vr, xt[:, chk])).max() / nc < 1e-5
assert np.abs(yb[:, chk] - orc.toeplitz_backward(vc, vr, xt[:, chk])).max() / nc < 1e-5
F = fm.Fourier(n)
nx = np.linalg.norm(x, axis=0).max() * 20
assert np.abs(F.forward(dev(x)).cpu().numpy()[:, chk] - orc.fourier_forward(x[:, chk])).max() / nx < 1e-5
print('switches ok', fm.launch_count())
'''


@pytest.mark.parametrize('env', [{'FMB_V32T': '1'}, {'FMB_V32T': '0', 'FMB_V32_PRUNE': '0'}, {'FMB_V32_TWM': '1', 'FMB_V32P': '2'},
                                 {'FMB_V32_OCC': '1', 'FMB_V32_MSHAPE': '3', 'FMB_V32P': '0'}, {'FMB_RM_CHUNK': '0', 'FMB_NO_V32': '1'},
                                 {'FMB_V32P_INPLACE': '0', 'FMB_V32_INPLACE': '0', 'FMB_FAST_INPLACE': '0'},
                                 {'FMB_V32P': '0', 'FMB_NO_V32': '1', 'FMB_FAST_INPLACE': '1'}],
                         ids=lambda e: ','.join('%s=%s' % kv for kv in e.items()))
def test_runtime_switches_keep_results_within_tolerance(fm, env):
    """The A/B switches of the 2^20 kernels (TMA-fed passes, pruning, twiddle placement, tile shapes,
    persistent kernel for convolutions, generic row-major route, 16-value fast path) all compute the same operator:
    Circulant / Toeplitz / Fourier at the BASELINE sizes against the oracle, both layouts."""
    out = _run_with_env(env, _SWITCH_CHECK)
    assert 'switches ok' in out


# ------------------------------------------------------------------------------------------- pipelined-slab schedule
def test_pipelined_slabs_match_single_columns(fm):
    """Column batches large enough for the pipelined-slab schedule (slabs round-robin on internal streams, ring of
    workspace slots, ragged last slab) must give bit-identical columns to one-column applies (single stream, no ring):
    the per-column arithmetic does not depend on the schedule."""
    n, m = 2 ** 20, 13
    rng = np.random.default_rng(99)
    g = torch.Generator(device='cuda').manual_seed(77)
    x = torch.complex(torch.randn((m, n), device='cuda', generator=g), torch.randn((m, n), device='cuda', generator=g)).t()
    c = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    ops = [fm.Circulant(c), fm.Fourier(n), fm.Kron(fm.Fourier(1024), fm.Fourier(1024))]
    for M in ops:
        for fn in (M.forward, M.backward):
            full = fn(x)
            for j in (0, 5, 11, 12):
                one = fn(x[:, j].contiguous())
                assert torch.equal(full[:, j], one), (repr(M), j)
    nt = n // 2
    T = fm.Toeplitz((rng.standard_normal(nt) + 1j * rng.standard_normal(nt)).astype(np.complex64),
                    (rng.standard_normal(nt - 1) + 1j * rng.standard_normal(nt - 1)).astype(np.complex64))
    xt = x[:nt].t().contiguous().t()
    full = T.forward(xt)
    for j in (0, 12):
        assert torch.equal(full[:, j], T.forward(xt[:, j].contiguous()))
    # FWHT: order 20, 27 float32 columns (6 slabs of 4 + 3)
    H = fm.Hadamard(20)
    xf = torch.randn((27, n), device='cuda', generator=g).t()
    full = H.forward(xf)
    for j in (0, 13, 26):
        assert torch.equal(full[:, j], H.forward(xf[:, j].contiguous()))
    # the caller's stream sees plain stream order: a dependent op right after the apply reads finished data
    y1 = ops[0].forward(x)
    s1 = y1.abs().sum()
    y2 = ops[0].forward(x)
    torch.cuda.synchronize()
    assert float(s1) == float(y2.abs().sum())


def test_cuda_graph_capture_of_pipelined_apply(fm):
    """The apply path issues only kernels and event fork/joins on streams (no allocation, no host sync), so a whole
    forward() - including the pipelined-slab schedule on the library's internal streams - can be captured in a CUDA graph
    and replayed on new data."""
    n, m = 2 ** 16, 200
    rng = np.random.default_rng(3)
    c = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    C = fm.Circulant(c)
    g = torch.Generator(device='cuda').manual_seed(8)
    x = torch.complex(torch.randn((m, n), device='cuda', generator=g), torch.randn((m, n), device='cuda', generator=g)).t()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                       # warm-up on the capture stream (first-use initialisation)
        for _ in range(2):
            y = C.forward(x)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        y = C.forward(x)
    eager = C.forward(x).clone()
    x.mul_(2.0)                                         # new data in the captured input buffer
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y, eager * 2.0) or float((y - eager * 2.0).abs().max()) <= 1e-6 * float(eager.abs().max())


# ------------------------------------------------------------------------------------------- persistent TMA pipeline (V32P)
_V32P_DUMP = r'''
import sys, numpy as np, torch
sys.path.insert(0, '.')
import fastmat_b200 as fm
N = 1 << 20
rng = np.random.default_rng(0)
g = torch.Generator(device='cuda').manual_seed(11)
c = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
def crandn(m, n):
    return torch.complex(torch.randn((m, n), device='cuda', generator=g), torch.randn((m, n), device='cuda', generator=g)).t()
xs = crandn(11, N)                                   # ragged last slab for 2-column slabs
C, F = fm.Circulant(c), fm.Fourier(N)
K = fm.Kron(fm.Fourier(1024), fm.Fourier(1024))
nt = N // 2
T = fm.Toeplitz(c[:nt].copy(), c[nt:2 * nt - 1].copy())
xt = crandn(5, nt)
res = {}
l0 = fm.launch_count()
for name, fn, x in (('circ_f', C.forward, xs), ('circ_b', C.backward, xs), ('four_f', F.forward, xs), ('four_b', F.backward, xs),
                    ('toep_f', T.forward, xt), ('toep_b', T.backward, xt), ('four_1', F.forward, xs[:, 3].contiguous()),
                    ('circ_view', C.forward, xs[:, 2:9]), ('kron_f', K.forward, xs), ('kron_b', K.backward, xs)):
    res[name] = fn(x).cpu()
res['launches'] = fm.launch_count() - l0
torch.save(res, sys.argv[1])
print('dumped')
'''


def test_v32p_persistent_tma_pipeline_bit_identical_to_per_pass_kernels(fm, tmp_path):
    """fft_v32p.cuh (one persistent launch, tiles requested by TMA ahead of the arithmetic, passes chained through global
    completion counters) executes the same butterflies in the same order per column as the per-pass V32 kernels: outputs
    must be BIT-identical for plain transforms (default path) and, opted in with FMB_V32P=2, for convolutions - including
    a ragged last slab, zero padding done by out-of-bounds rows of the tensor map (Toeplitz), single columns, column
    views - and the whole apply is one kernel launch."""
    import subprocess
    import sys
    from conftest import ROOT
    outs = {}
    for mode in ('0', '1', '2'):
        path = str(tmp_path / ('v32p_%s.pt' % mode))
        e = dict(os.environ)
        e['FMB_V32P'] = mode
        # same arithmetic on both sides: the per-pass convolution applies the inverse transform's four-step twiddle on the
        # last pass's loads by default, the persistent kernel on the middle pass's stores (= FMB_V32_TWM=1)
        e['FMB_V32_TWM'] = '1'
        e['FMB_V32_PRUNE'] = '0'                 # ... and the persistent kernel runs the full butterflies on zero-padded halves
        r = subprocess.run([sys.executable, '-c', _V32P_DUMP, path], cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        outs[mode] = torch.load(path)
    for mode in ('1', '2'):
        for k, ref in outs['0'].items():
            if k != 'launches':
                assert torch.equal(ref, outs[mode][k]), (mode, k)
    # per-pass: 3 (2) launches per slab; persistent: one launch per plain transform (mode 1), one per apply (mode 2)
    assert outs['2']['launches'] <= 12, outs['2']['launches']          # 10 applies (+ a layout copy for the column view)
    assert outs['0']['launches'] > outs['1']['launches'] > outs['2']['launches']


def test_v32p_apply_in_cuda_graph(fm):
    """The persistent launch (memset of the completion counters + one kernel taking two tensor maps by value) can be
    captured in a CUDA graph and replayed on new data."""
    n, m = 2 ** 20, 6
    F = fm.Fourier(n)
    g = torch.Generator(device='cuda').manual_seed(8)
    x = torch.complex(torch.randn((m, n), device='cuda', generator=g), torch.randn((m, n), device='cuda', generator=g)).t()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            y = F.forward(x)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        y = F.forward(x)
    eager = F.forward(x).clone()
    x.mul_(2.0)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(y, eager * 2.0)


# ------------------------------------------------------------------------------------------- round-2 advisor items
@pytest.mark.gpu
def test_dense_matrix_uses_numpy_promotion_not_torch(fm):
    """fastmat's promotion table is np.promote_types (fastmat/core/types.pyx:443-453): an int32 / int64 matrix applied to
    float32 data is float64 (torch.promote_types would say float32 and lose integers above 2^24), integer x integer wraps
    like numpy.  Blocks accumulates its terms in the same table."""
    rng = np.random.default_rng(5)
    a = rng.integers(2 ** 24, 2 ** 26, size=(6, 5)).astype(np.int32)
    x = rng.integers(1, 4, size=(5, 3)).astype(np.float32)
    M = fm.Matrix(torch.from_numpy(a))
    y = M.forward(torch.from_numpy(x).cuda())
    ref = a.astype(np.float64) @ x.astype(np.float64)
    assert y.dtype == torch.float64 and np.array_equal(y.cpu().numpy(), ref)
    yb = M.backward(torch.from_numpy(ref.astype(np.float32)).cuda())
    assert yb.dtype == torch.float64
    xi = rng.integers(-2 ** 40, 2 ** 40, size=(5, 2)).astype(np.int64)
    a64 = rng.integers(-2 ** 40, 2 ** 40, size=(4, 5)).astype(np.int64)
    yi = fm.Matrix(torch.from_numpy(a64)).forward(torch.from_numpy(xi).cuda())
    assert yi.dtype == torch.int64 and np.array_equal(yi.cpu().numpy(), a64 @ xi)          # wraps exactly like numpy
    B = fm.Blocks([[fm.Matrix(torch.from_numpy(a)), fm.Matrix(torch.from_numpy(a))]])
    xb = torch.from_numpy(np.vstack([x, x])).cuda()
    yB = B.forward(xb)
    assert yB.dtype == torch.float64 and np.array_equal(yB.cpu().numpy(), 2 * ref)


@pytest.mark.gpu
def test_product_python_scalars_are_weakly_typed(fm):
    """The operator dtype never depends on a scalar's VALUE (NEP 50 semantics): M * 0.5 and M * 0.1, M / 2 and M / 3 have the
    same dtype; a python scalar lifts integer operators to the reference's int64 / float64 / complex128
    (np.array(scalar).dtype, fastmat/Product.pyx:108-114) and leaves the precision of floating operators alone; numpy
    scalars keep their dtype."""
    rng = np.random.default_rng(6)
    c = (rng.standard_normal(16) + 1j * rng.standard_normal(16)).astype(np.complex64)
    C = fm.Circulant(c)
    assert (C * 0.5).dtype == (C * 0.1).dtype == (C / 2).dtype == (C / 3).dtype == np.complex64
    assert (2 * C).dtype == np.complex64 and (C * 2j).dtype == np.complex64
    assert (C * np.float64(0.5)).dtype == np.complex128
    H = fm.Hadamard(3)
    # integer operators: python int -> int64, then Product's safe type expansion (types.pyx:378-394) makes it float64
    assert (H * 2).dtype == (H / 2).dtype == (H / 3).dtype == np.float64 and (H * 1j).dtype == np.complex128
    assert (H * np.int8(2)).dtype == np.float32                                  # int8 x int8 -> int8 -> expanded to float32
    assert fm.Product(H, 2, typeExpansion=None).dtype == np.int64 and fm.Product(H, np.int8(2), typeExpansion=None).dtype == np.int8
    D = fm.Diag(np.arange(1, 9, dtype=np.float32))
    assert (D * 0.1).dtype == np.float32 and (D * 1j).dtype == np.complex64
    x = torch.from_numpy(rng.standard_normal((16, 2)).astype(np.float32)).cuda().to(torch.complex64)
    y1, y2 = (C / 3).forward(x), C.forward(x) / 3
    assert y1.dtype == torch.complex64 and float((y1 - y2).abs().max()) <= 1e-5 * float(y2.abs().max())


@pytest.mark.gpu
def test_toeplitz_norms_closed_form(fm):
    """Toeplitz colNorms / rowNorms (fastmat/Toeplitz.pyx:371-625) from cumulative sums of |t|^2, single- and multi-level,
    rectangular levels included, against the dense matrix by index placement; usable at 2^19 x 2^19 (O(n), no applies)."""
    rng = np.random.default_rng(31)
    for dr, dc in (((5, ), (3, )), ((3, ), (6, )), ((1, ), (5, )), ((3, 4), (2, 5)), ((2, 3, 2), (3, 2, 4))):
        shape = tuple(a + b - 1 for a, b in zip(dr, dc))
        t = (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex128)
        T = fm.Toeplitz(t, split=list(dr))
        D = T.reference().cpu().numpy()
        assert D.shape == (int(np.prod(dr)), int(np.prod(dc)))
        assert np.abs(T.colNorms.cpu().numpy() - np.linalg.norm(D, axis=0)).max() <= 1e-12 * np.linalg.norm(D)
        assert np.abs(T.rowNorms.cpu().numpy() - np.linalg.norm(D, axis=1)).max() <= 1e-12 * np.linalg.norm(D)
    n = 1 << 19
    vc, vr = seeded(61, n).astype(np.complex64), seeded(62, n - 1).astype(np.complex64)
    T = fm.Toeplitz(vc, vr)
    launches = fm.launch_count()
    cn = T.colNorms.cpu().numpy()
    assert fm.launch_count() == launches                                   # closed form: not a single apply
    a2, r2 = np.abs(vc.astype(np.complex128)) ** 2, np.abs(vr.astype(np.complex128)) ** 2
    assert abs(cn[0] ** 2 - a2.sum()) <= 1e-9 * a2.sum()                    # column 0 is vecC
    assert abs(cn[n - 1] ** 2 - (a2[0] + r2.sum())) <= 1e-9 * r2.sum()      # last column: t[0] and the whole row part
    x = dev(seeded(63, n, 2).astype(np.complex64))
    y = T.colNormalized.forward(x)                                          # usable at the BASELINE size now
    ref = T.forward(x / torch.from_numpy(cn).cuda().reshape(-1, 1).to(torch.complex64))
    assert float((y - ref).abs().max()) <= 1e-4 * float(ref.abs().max())


@pytest.mark.gpu
def test_largest_eigenvalue_and_scipy_linear_operator(fm):
    """largestEigenValue (fastmat/Matrix.pyx:678-760) by power iteration on the device; scipyLinearOperator
    (:977-1006) drives scipy's Krylov solvers with host vectors through apply_host."""
    from scipy.sparse.linalg import svds
    rng = np.random.default_rng(41)
    ev = 0.8 * np.exp(2j * np.pi * rng.random(64)) * rng.random(64)         # eigenvalues of a circulant = fft of its first column
    ev[3] = 5.0 * np.exp(0.7j)                                              # one dominant, well separated (complex) eigenvalue
    C = fm.Circulant(np.fft.ifft(ev).astype(np.complex128))
    assert abs(C.largestEigenValue - ev[3]) <= 1e-8 * abs(ev[3])
    with pytest.raises(ValueError):
        fm.Partial(fm.Fourier(8), rows=np.arange(4)).largestEigenValue
    A = fm.Product(fm.Partial(fm.Fourier(64), rows=np.arange(0, 64, 3)), fm.Diag(rng.standard_normal(64) + 2.0))
    op = A.scipyLinearOperator
    assert op.shape == (22, 64)
    x = rng.standard_normal(64) + 1j * rng.standard_normal(64)
    assert np.abs(op.matvec(x) - A.forward(torch.from_numpy(x).cuda()).cpu().numpy()).max() <= 1e-10
    s = svds(op, k=1, return_singular_vectors=False)[0]
    assert abs(s - A.largestSingularValue) <= 1e-6 * s


@pytest.mark.gpu
def test_input_on_another_device_is_rejected(fm):
    """A matrix belongs to the device that was current when it was built (plan constants live there)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    F = fm.Fourier(16)
    with pytest.raises(RuntimeError):
        F.forward(torch.zeros(16, dtype=torch.complex64, device='cuda:1'))


@pytest.mark.gpu
def test_concurrent_host_threads_on_one_plan(fm):
    """include/fastmat_b200.h: apply is safe to call concurrently on one plan from several host threads / streams with
    distinct workspaces.  Each host thread has its own internal streams for the pipelined-slab schedule (no issue mutex);
    four threads x three applies of one Circulant 2^20 plan on their own streams give the serial results bit for bit."""
    import threading
    n, m = 2 ** 20, 13
    rng = np.random.default_rng(123)
    C = fm.Circulant((rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64))
    g = torch.Generator(device='cuda').manual_seed(5)
    xs = [torch.complex(torch.randn((m, n), device='cuda', generator=g), torch.randn((m, n), device='cuda', generator=g)).t()
          for _ in range(4)]
    ref = [C.forward(x) for x in xs]
    torch.cuda.synchronize()
    out, errs = [None] * 4, []

    def work(i):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(3):
                    y = C.forward(xs[i])
            s.synchronize()
            out[i] = y
        except Exception as e:                                   # noqa: BLE001
            errs.append(repr(e))

    threads = [threading.Thread(target=work, args=(i, )) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for i in range(4):
        assert torch.equal(out[i], ref[i]), i


@pytest.mark.gpu
@pytest.mark.parametrize('n', [2 ** 13, 2 ** 14, 2 ** 15])
def test_pass_length_128_fast_path(fm, orc, n):
    """2^13 = 64 x 128 (FastPlan<6>: radix 16 then radix 4), 2^14 = 128 x 128 and 2^15 = 128 x 256 run the compile-time specialised passes since round 2 (FastPlan<7>: radix 16
    then radix 8); round 1 sent them to the generic run-time-radix kernel.  Fourier / Circulant / Toeplitz, forward and
    backward, both precisions, against the oracle."""
    x = seeded(700 + n % 97, n, 5)
    c = seeded(701, n)
    for dt, tol in ((np.complex64, TOL64), (np.complex128, TOL128)):
        xd = dev(x.astype(dt))
        nx = np.linalg.norm(x, axis=0).max() * np.log2(n)
        F = fm.Fourier(n)
        assert np.abs(F.forward(xd).cpu().numpy() - orc.fourier_forward(x)).max() / nx < tol
        assert np.abs(F.backward(xd).cpu().numpy() - orc.fourier_backward(x)).max() / nx < tol
        C = fm.Circulant(c.astype(dt))
        ncn = np.linalg.norm(c) * nx
        assert np.abs(C.forward(xd).cpu().numpy() - orc.circulant_forward(c, x)).max() / ncn < tol
        assert np.abs(C.backward(xd).cpu().numpy() - orc.circulant_backward(c, x)).max() / ncn < tol
    nt = n // 2
    T = fm.Toeplitz(c[:nt].astype(np.complex64), c[nt:2 * nt - 1].astype(np.complex64))
    xt = x[:nt]
    ntn = np.linalg.norm(c) * np.linalg.norm(xt, axis=0).max() * np.log2(n)
    assert np.abs(T.forward(dev(xt.astype(np.complex64))).cpu().numpy() - orc.toeplitz_forward(c[:nt], c[nt:2 * nt - 1], xt)).max() / ntn < TOL64
    assert np.abs(T.backward(dev(xt.astype(np.complex64))).cpu().numpy() - orc.toeplitz_backward(c[:nt], c[nt:2 * nt - 1], xt)).max() / ntn < TOL64


@pytest.mark.gpu
@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096])
def test_single_kernel_fast_path(fm, orc, n):
    """Power-of-two lengths 128 ... 4096 run as ONE specialised kernel per apply (fft_engine.cu: run_single_fast): a line
    of the pass kernel is a column of the operand; Circulant / Toeplitz do FFT -> spectrum -> FFT on chip with the zero
    padding as load mask.  Column counts that are not a multiple of the tile (the tail takes the generic kernel), a
    strided (non-contiguous) column batch, both precisions and both directions, against the oracle."""
    cols = 3 * 64 + 5
    x = seeded(900 + n % 89, n, cols)
    c = seeded(901, n)
    for dt, tol in ((np.complex64, TOL64), (np.complex128, TOL128)):
        xd = dev(x.astype(dt))
        nx = np.linalg.norm(x, axis=0).max() * np.log2(n)
        F = fm.Fourier(n)
        assert np.abs(F.forward(xd).cpu().numpy() - orc.fourier_forward(x)).max() / nx < tol
        assert np.abs(F.backward(xd).cpu().numpy() - orc.fourier_backward(x)).max() / nx < tol
        C = fm.Circulant(c.astype(dt))
        ncn = np.linalg.norm(c) * nx
        assert np.abs(C.forward(xd).cpu().numpy() - orc.circulant_forward(c, x)).max() / ncn < tol
        assert np.abs(C.backward(xd).cpu().numpy() - orc.circulant_backward(c, x)).max() / ncn < tol
        # every second column of a wider column-major batch: column stride 2 n
        wide = dev(np.repeat(x.astype(dt), 2, axis=1))
        assert np.abs(F.forward(wide[:, ::2]).cpu().numpy() - orc.fourier_forward(x)).max() / nx < tol
        # row-major operands (torch's default layout): the line-fastest variants of the same kernels
        xr = torch.from_numpy(np.ascontiguousarray(x.astype(dt))).cuda()
        assert xr.stride(1) == 1
        assert np.abs(F.forward(xr).cpu().numpy() - orc.fourier_forward(x)).max() / nx < tol
        assert np.abs(F.backward(xr).cpu().numpy() - orc.fourier_backward(x)).max() / nx < tol
        assert np.abs(C.forward(xr).cpu().numpy() - orc.circulant_forward(c, x)).max() / ncn < tol
        assert np.abs(C.backward(xr).cpu().numpy() - orc.circulant_backward(c, x)).max() / ncn < tol
        # Toeplitz of order n/2 + 1 ... pads to n: rows beyond the order are masked on load and cropped on store
        nt = n // 2
        T = fm.Toeplitz(c[:nt].astype(dt), c[nt:2 * nt - 1].astype(dt))
        xt = x[:nt]
        ntn = np.linalg.norm(c) * np.linalg.norm(xt, axis=0).max() * np.log2(n)
        assert np.abs(T.forward(dev(xt.astype(dt))).cpu().numpy() - orc.toeplitz_forward(c[:nt], c[nt:2 * nt - 1], xt)).max() / ntn < tol
        assert np.abs(T.backward(dev(xt.astype(dt))).cpu().numpy() - orc.toeplitz_backward(c[:nt], c[nt:2 * nt - 1], xt)).max() / ntn < tol
        xtr = torch.from_numpy(np.ascontiguousarray(xt.astype(dt))).cuda()
        assert np.abs(T.forward(xtr).cpu().numpy() - orc.toeplitz_forward(c[:nt], c[nt:2 * nt - 1], xt)).max() / ntn < tol
        assert np.abs(T.backward(xtr).cpu().numpy() - orc.toeplitz_backward(c[:nt], c[nt:2 * nt - 1], xt)).max() / ntn < tol
    if n == 1024:
        # FFT length 1024 runs the 32-value kernel: Circulant above, the pruned Toeplitz (512 x 512) above, and here a
        # non-square Toeplitz (500 x 512 padded to 1024: masked loads and stores) and a strided column batch
        vc, vr = c[:500].astype(np.complex64), c[500:1011].astype(np.complex64)
        T2 = fm.Toeplitz(vc, vr)
        assert T2.shape == (500, 512)
        x2 = x[:512]
        n2 = np.linalg.norm(c) * np.linalg.norm(x2, axis=0).max() * np.log2(n)
        assert np.abs(T2.forward(dev(x2.astype(np.complex64))).cpu().numpy() - orc.toeplitz_forward(vc, vr, x2)).max() / n2 < TOL64
        y2 = x[:500]
        assert np.abs(T2.backward(dev(y2.astype(np.complex64))).cpu().numpy() - orc.toeplitz_backward(vc, vr, y2)).max() / n2 < TOL64
        C64 = fm.Circulant(c.astype(np.complex64))
        wide = dev(np.repeat(x.astype(np.complex64), 2, axis=1))
        ncn = np.linalg.norm(c) * np.linalg.norm(x, axis=0).max() * np.log2(n)
        assert np.abs(C64.forward(wide[:, ::2]).cpu().numpy() - orc.circulant_forward(c, x)).max() / ncn < TOL64
    # the two routes agree to rounding on the same input (same algorithm, different radix plan)
    import os
    import subprocess
    import sys
    code = ("import numpy as np, torch, fastmat_b200 as fm\n"
            "rng = np.random.default_rng(5); n = %d\n"
            "x = torch.from_numpy((rng.standard_normal((64, n)) + 1j * rng.standard_normal((64, n))).astype(np.complex64)).cuda().t()\n"
            "y = fm.Fourier(n).forward(x)\n"
            "print(float(torch.linalg.vector_norm(y - torch.fft.fft(x.to(torch.complex128), dim=0).to(torch.complex64)) / torch.linalg.vector_norm(y)))\n" % n)
    errs = []
    for flag in ("0", "1"):
        env = dict(os.environ, FMB_NO_FAST1=flag)
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300,
                             cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert out.returncode == 0, out.stderr[-2000:]
        errs.append(float(out.stdout.strip().splitlines()[-1]))
    assert max(errs) < 1e-6, errs


@pytest.mark.gpu
@pytest.mark.parametrize("dims", [(64, 64), (64, 128), (256, 64)])
def test_kron_fourier_pass_length_64(fm, orc, dims):
    """Kron(Fourier(a), Fourier(b)) with a factor of 64 runs the specialised passes (FastPlan<6>) instead of the generic
    kernel; forward and backward, both precisions, against the N-D FFT of the oracle."""
    n = dims[0] * dims[1]
    x = seeded(950 + dims[0], n, 7)
    K = fm.Kron(fm.Fourier(dims[0]), fm.Fourier(dims[1]))
    nx = np.linalg.norm(x, axis=0).max() * np.log2(n)
    for dt, tol in ((np.complex64, TOL64), (np.complex128, TOL128)):
        xd = dev(x.astype(dt))
        assert np.abs(K.forward(xd).cpu().numpy() - orc.kron_fourier_forward(dims, x)).max() / nx < tol
        assert np.abs(K.backward(xd).cpu().numpy() - orc.kron_fourier_backward(dims, x)).max() / nx < tol


@pytest.mark.gpu
@pytest.mark.parametrize("order", [24, 28])
def test_hadamard_large_orders_exact(fm, order):
    """Orders above 20 chain the 12-bit first pass (64 values per thread) with strided passes: 12 + 6 + 6 for order 24,
    12 + 8 + 8 for order 28 (the 128-bit strided kernel with a compile-time and with a run-time stride).  int32 ring
    arithmetic makes everything exact: H(H x) = 2^order x, and single outputs against the defining signed sum."""
    n = 2 ** order
    H = fm.Hadamard(order)
    g = torch.Generator(device='cuda').manual_seed(order)
    x = torch.randint(-2 ** 31, 2 ** 31 - 1, (1, n), dtype=torch.int32, device='cuda', generator=g).t()
    y = H.forward(x)
    idx = torch.arange(n, device='cuda', dtype=torch.int64)
    for k in (0, 1, n - 1, 0x5A5A5A5 % n, 3 * 4096 + 17, (1 << (order - 1)) + 4096 * 255 + 77):
        v = idx & k                                       # parity of popcount(n & k) by xor folding
        for sh in (32, 16, 8, 4, 2, 1):
            v = v ^ (v >> sh)
        sign = 1 - 2 * (v & 1)
        want = int((x[:, 0].to(torch.int64) * sign).sum().item())
        want = (want + 2 ** 31) % 2 ** 32 - 2 ** 31       # wrap to int32
        assert int(y[k, 0].item()) == want, k
    z = H.forward(y)
    assert torch.equal(z, x * (2 ** order))


@pytest.mark.gpu
def test_hadamard_unaligned_columns_bit_identical(fm):
    """The bulk-copy / 128-bit kernels need 16-byte aligned columns; a batch whose column stride is odd takes the scalar
    kernels and must give the same bits (order 20, float32 and int32)."""
    order, m = 20, 5
    n = 2 ** order
    H = fm.Hadamard(order)
    g = torch.Generator(device='cuda').manual_seed(11)
    for dt in (torch.float32, torch.int32):
        if dt == torch.float32:
            base = torch.randn((m, n + 1), dtype=dt, device='cuda', generator=g)
        else:
            base = torch.randint(-2 ** 31, 2 ** 31 - 1, (m, n + 1), dtype=dt, device='cuda', generator=g)
        xu = base[:, 1:].t()                              # column stride n + 1, first element 4 bytes off alignment
        assert xu.stride(0) == 1 and xu.stride(1) == n + 1
        xa = xu.t().contiguous().t()
        assert torch.equal(H.forward(xu), H.forward(xa))


_INPLACE_DIGEST = r'''
import sys, hashlib, numpy as np, torch
sys.path.insert(0, '.')
import fastmat_b200 as fm
g = torch.Generator(device='cuda').manual_seed(21)
def crandn(n, m): return torch.complex(torch.randn((m, n), device='cuda', generator=g), torch.randn((m, n), device='cuda', generator=g)).t()
h = hashlib.sha256()
n, m = 2 ** 20, 13
x = crandn(n, m)
c = np.random.default_rng(4).standard_normal(n).astype(np.complex64)
keep = x.clone()
for op in (fm.Fourier(n), fm.Kron(fm.Fourier(1024), fm.Fourier(1024)), fm.Circulant(c)):
    for fn in (op.forward, op.backward):
        y = fn(x)
        assert torch.equal(x, keep)                       # the intermediate lives in y, never in x
        h.update(torch.view_as_real(y).cpu().numpy().tobytes())
x2 = crandn(2 ** 14, 300)
h.update(torch.view_as_real(fm.Fourier(2 ** 14).forward(x2)).cpu().numpy().tobytes())
print('digest', h.hexdigest())
'''


@pytest.mark.gpu
def test_in_place_intermediate_is_bit_identical_to_ring(fm):
    """Round 2 keeps the intermediate of plain transforms (persistent 2^20 kernel, 16-value two-pass path) and of the
    Circulant in y instead of a ring in the workspace (FMB_V32P_INPLACE / FMB_V32_INPLACE / FMB_FAST_INPLACE).  Only the
    location of the intermediate changes: outputs must be bit-identical to the ring variants, forward and backward, and
    the input must stay untouched."""
    a = _run_with_env({}, _INPLACE_DIGEST)
    b = _run_with_env({'FMB_V32P_INPLACE': '0', 'FMB_V32_INPLACE': '0', 'FMB_FAST_INPLACE': '0'}, _INPLACE_DIGEST)
    da = [ln for ln in a.splitlines() if ln.startswith('digest')]
    db = [ln for ln in b.splitlines() if ln.startswith('digest')]
    assert da and da == db, (da, db)


_FWHT_DIGEST = r'''
import sys, hashlib, torch
sys.path.insert(0, '.')
import fastmat_b200 as fm
g = torch.Generator(device='cuda').manual_seed(31)
H = fm.Hadamard(20)
h = hashlib.sha256()
for m in (37, 16, 130):
    xf = torch.randn((m, 2 ** 20), device='cuda', generator=g).t()
    xi = torch.randint(-2 ** 31, 2 ** 31 - 1, (m, 2 ** 20), dtype=torch.int32, device='cuda', generator=g).t()
    for x in (xf, xi):
        keep = x.clone()
        y = H.forward(x)
        assert torch.equal(x, keep)
        h.update(y.cpu().numpy().tobytes())
print('digest', h.hexdigest(), fm.launch_count())
'''


@pytest.mark.gpu
@pytest.mark.parametrize('env', [{'FMB_FWHT_PERSIST': '1'}, {'FMB_FWHT_PERSIST': '1', 'FMB_FWHT_PERSIST_NT': '128', 'FMB_FWHT_PERSIST_SLAB': '1', 'FMB_FWHT_PERSIST_DIST': '4'},
                                 {'FMB_FWHT_NO12': '1', 'FMB_FWHT_NO_S8': '1'}, {'FMB_FWHT_PIPE_STREAMS': '1'}],
                         ids=lambda e: ','.join('%s=%s' % kv for kv in e.items()))
def test_fwht_schedules_are_bit_identical(fm, env):
    """Order-20 FWHT, float32 and int32, ragged column counts: the persistent cooperative kernel (opt-in experiment, both
    CTA shapes), the round-1 pass kernels and the single-stream schedule all produce the bits of the default path (same
    butterflies in the same order; only the schedule and the data movement differ)."""
    a = [ln.split()[1] for ln in _run_with_env({}, _FWHT_DIGEST).splitlines() if ln.startswith('digest')]
    b = [ln.split()[1] for ln in _run_with_env(env, _FWHT_DIGEST).splitlines() if ln.startswith('digest')]
    assert a and a == b, (a, b)


_REAL_CHECK = r'''
import sys, numpy as np, torch
sys.path.insert(0, '.')
import fastmat_b200 as fm
from oracle import fastmat_oracle as orc
rng = np.random.default_rng(12)
for n, m in ((2 ** 14, 21), (1024, 70), (2 ** 20, 3)):
    for dt, tol in ((np.float32, 1e-5), (np.float64, 1e-12)):
        x = rng.standard_normal((n, m)).astype(dt)
        c = rng.standard_normal(n).astype(dt)
        xd = torch.from_numpy(np.asfortranarray(x).T.copy()).cuda().t()
        nx = np.linalg.norm(x, axis=0).max() * np.log2(n)
        y = fm.Fourier(n).forward(xd)
        assert y.dtype == (torch.complex64 if dt == np.float32 else torch.complex128)
        assert np.abs(y.cpu().numpy() - orc.fourier_forward(x)).max() / nx < tol
        C = fm.Circulant(c)
        yc = C.forward(xd).cpu().numpy()
        assert np.abs(yc - orc.circulant_forward(c, x)).max() / (np.linalg.norm(c) * nx) < tol
        yb = C.backward(xd).cpu().numpy()
        assert np.abs(yb - orc.circulant_backward(c, x)).max() / (np.linalg.norm(c) * nx) < tol
        nt = n // 2
        T = fm.Toeplitz(c[:nt], c[nt:2 * nt - 1])
        yt = T.forward(xd[:nt]).cpu().numpy()
        assert np.abs(yt - orc.toeplitz_forward(c[:nt], c[nt:2 * nt - 1], x[:nt])).max() / (np.linalg.norm(c) * nx) < tol
print('real ok')
'''


@pytest.mark.gpu
@pytest.mark.parametrize('cast', ['1', '0'])
def test_real_operands_on_fft_operators(fm, cast):
    """float32 / float64 operands of Fourier / Circulant / Toeplitz at power-of-two sizes: widened to complex in front of
    the specialised kernels (default) or read directly by the generic real-input kernels (FMB_REAL_CAST=0); both against
    the oracle (fastmat returns complex results for real operands of these operators)."""
    assert 'real ok' in _run_with_env({'FMB_REAL_CAST': cast}, _REAL_CHECK)


@pytest.mark.gpu
def test_strided_output_through_the_c_abi_keeps_the_gaps(fm):
    """The in-place variants use y itself as the intermediate.  Through the C-ABI y may be a strided view (column stride
    larger than the row count): only rows [0, n) of every column may be written, the padding between the columns must keep
    its bytes, and the result must equal the contiguous call.  Fourier / Kron / Circulant 2^20 (persistent kernel and
    per-pass route), Fourier 2^14 (16-value path), Circulant 1024 (single kernel), Hadamard order 20 (in-place passes)."""
    from fastmat_b200 import _lib
    from fastmat_b200.Matrix import _stream_ptr
    from fastmat_b200.core import types as T
    rng = np.random.default_rng(17)
    g = torch.Generator(device='cuda').manual_seed(17)

    def crandn(n, m):
        return torch.complex(torch.randn((m, n), device='cuda', generator=g), torch.randn((m, n), device='cuda', generator=g)).t()

    def cvec(n):
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)

    cases = [(fm.Fourier(2 ** 20), crandn(2 ** 20, 9)), (fm.Kron(fm.Fourier(1024), fm.Fourier(1024)), crandn(2 ** 20, 9)),
             (fm.Circulant(cvec(2 ** 20)), crandn(2 ** 20, 13)), (fm.Fourier(2 ** 14), crandn(2 ** 14, 300)),
             (fm.Circulant(cvec(1024)), crandn(1024, 100)), (fm.Hadamard(20), torch.randn((27, 2 ** 20), device='cuda', generator=g).t())]
    for op, x in cases:
        n, m = x.shape
        pad = 64
        ref = op.forward(x)
        sentinel = 12345.0
        buf = torch.full((m, n + pad), sentinel, dtype=ref.dtype, device='cuda')
        y = buf[:, :n].t()                                 # column-major view, column stride n + pad
        assert y.stride(0) == 1 and y.stride(1) == n + pad
        ft = T.getFusedType(x.dtype)
        fo = T.getFusedType(ref.dtype)
        plan = op._plan
        wsb = _lib.lib.fmb_plan_workspace_bytes(plan.handle, _lib.FORWARD, m, ft, fo)
        ws = torch.empty(max(wsb, 1), dtype=torch.uint8, device='cuda')
        _lib.check(_lib.lib.fmb_plan_apply(plan.handle, _lib.FORWARD, x.data_ptr(), x.stride(0), x.stride(1), y.data_ptr(), y.stride(0),
                                           y.stride(1), m, ft, fo, ws.data_ptr(), wsb, _stream_ptr(x.device)))
        torch.cuda.synchronize()
        assert torch.equal(y, ref), repr(op)
        assert bool((buf[:, n:] == sentinel).all()), repr(op)


_PAD_CHECK = r'''
import sys, numpy as np, torch
sys.path.insert(0, '.')
import fastmat_b200 as fm
from oracle import fastmat_oracle as orc
rng = np.random.default_rng(23)
def dev(a): return torch.from_numpy(np.ascontiguousarray(a.T)).cuda().t()
inner = []
for n in (41, 100, 1000, 1500, 3000, 6144, 5000, 41000):
    for dt, tol in ((np.complex64, 1e-5), (np.complex128, 1e-12)):
        x = (rng.standard_normal((n, 9)) + 1j * rng.standard_normal((n, 9))).astype(dt)
        c = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(dt)
        nx = np.linalg.norm(x, axis=0).max() * np.log2(n)
        C = fm.Circulant(c)
        assert np.abs(C.forward(dev(x)).cpu().numpy() - orc.circulant_forward(c, x)).max() / (np.linalg.norm(c) * nx) < tol
        assert np.abs(C.backward(dev(x)).cpu().numpy() - orc.circulant_backward(c, x)).max() / (np.linalg.norm(c) * nx) < tol
        vr = c[1:][::-1].copy()
        T = fm.Toeplitz(c, vr)
        assert np.abs(T.forward(dev(x)).cpu().numpy() - orc.toeplitz_forward(c, vr, x)).max() / (np.linalg.norm(c) * nx) < tol
        assert np.abs(T.backward(dev(x)).cpu().numpy() - orc.toeplitz_backward(c, vr, x)).max() / (np.linalg.norm(c) * nx) < tol
        F = fm.Fourier(n)
        assert np.abs(F.forward(dev(x)).cpu().numpy() - orc.fourier_forward(x)).max() / nx < tol
        assert np.abs(F.backward(dev(x)).cpu().numpy() - orc.fourier_backward(x)).max() / nx < tol
        xr = torch.from_numpy(np.ascontiguousarray(x)).cuda()          # row-major operand
        assert np.abs(F.forward(xr).cpu().numpy() - orc.fourier_forward(x)).max() / nx < tol
        xw = (rng.standard_normal((n, 70)) + 1j * rng.standard_normal((n, 70))).astype(dt)   # whole tiles + ragged tail
        assert np.abs(F.forward(dev(xw)).cpu().numpy() - orc.fourier_forward(xw)).max() / (np.linalg.norm(xw, axis=0).max() * np.log2(n)) < tol
    inner.append((int(C._plan.info.inner_size), int(T._plan.info.inner_size), int(F._plan.info.inner_size)))
print('pad ok', inner)
'''


@pytest.mark.gpu
@pytest.mark.parametrize('pad', ['1', '0'])
def test_padded_length_policy(fm, pad):
    """Lengths that are not powers of two: by default Circulant / Toeplitz pad to the next power of two that embeds them
    and Fourier orders above 4096 run as chirp-z transforms over a power of two (specialised kernels); FMB_POW2_PAD=0 keeps
    the reference planner's choice (2^a 3^b 5^c lengths, run-time-radix kernels).  Same operator either way: both against
    the oracle, both precisions, forward and backward."""
    out = _run_with_env({'FMB_POW2_PAD': pad}, _PAD_CHECK)
    assert 'pad ok' in out
    inner = eval(out.split('pad ok', 1)[1].strip().splitlines()[0])
    pow2 = [all(v & (v - 1) == 0 for v in t[:2]) for t in inner]
    assert all(pow2) if pad == '1' else not all(pow2)
