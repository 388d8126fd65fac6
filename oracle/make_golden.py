#!/usr/bin/env python
"""Generate tests/golden/* from the REAL reference (TEST INFRASTRUCTURE ONLY).

Run in the build container (needs oracle/_ref, built by oracle/build_ref.sh from /root/reference):

    bash oracle/build_ref.sh && python oracle/make_golden.py

The reference has no stored golden vectors (its tests draw unseeded random inputs and compare against dense
``reference()`` matrices, fastmat/inspect/test.py:318-368), so the fixtures frozen here are seeded runs of the
reference itself on the hot-path shapes its own test-suite uses (SURVEY.md section 4 "hot-path test vectors")
plus the planner known-answer tables.  Floating-point cases are evaluated in double
(inputs and generators cast to complex128/float64 first: SURVEY.md section 8c rule 1); the as-is dtype behaviour of
the reference is recorded separately in ``dtypes``.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '_ref'))
import fastmat as fm                                    # noqa: E402  (the real reference)
from fastmat.core.cmath import _findOptimalFFTSize, _getFFTComplexity  # noqa: E402

OUT = os.path.join(HERE, '..', 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)
rng = np.random.default_rng(20261017)
ALL_TYPES = ['int8', 'int16', 'int32', 'int64', 'float32', 'float64', 'complex64', 'complex128']

arrays = {}
meta = {'cases': {}, 'planner': {}, 'dtypes': {}, 'versions': {
    'fastmat': fm.__version__, 'numpy': np.__version__}}


def crand(*shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def seeded(seed, *shape):
    """Large inputs are NOT stored: tests regenerate them from the seed with this exact recipe."""
    g = np.random.default_rng(seed)
    return g.standard_normal(shape) + 1j * g.standard_normal(shape)


def put_big(name, kind, params, **outs):
    """Store outputs of a large case at 512 pseudo-random rows only (plus the full-array sum as a checksum)."""
    arrs = {}
    for k, v in outs.items():
        n = v.shape[0]
        rows = np.sort(np.random.default_rng(n).choice(n, size=min(n, 512), replace=False))
        arrs[k + '_rows'] = rows
        arrs[k] = v[rows]
        arrs[k + '_sum'] = np.asarray(v.sum(axis=0))
    put(name, kind, params, **arrs)


def put(name, kind, params, **arrs):
    meta['cases'][name] = {'kind': kind, 'params': params, 'arrays': sorted(arrs)}
    for k, v in arrs.items():
        arrays['%s/%s' % (name, k)] = np.asarray(v)


# ------------------------------------------------------------------ planner known answers (bit-exact ints/floats)
orders = sorted(set(list(range(1, 300)) + [2 ** k for k in range(1, 25)] + [2 ** k - 1 for k in range(2, 25)] +
                    [2 ** k + 1 for k in range(2, 25)] +
                    [1000, 1001, 1021, 4093, 4097, 65537, 100000, 1000003, 2000005, 3 * 2 ** 18, 5 ** 7, 7 ** 6]))
meta['planner']['orders'] = orders
for ms in (2, 3, 4, 5, 7):
    meta['planner']['opt_%d' % ms] = [int(_findOptimalFFTSize(o, ms)) for o in orders]
meta['planner']['complexity_bits'] = [int(np.float32(_getFFTComplexity(o)).view(np.uint32)) for o in orders]
small = list(range(1, 400)) + [509, 1000, 1001, 1021, 4093, 65537, 2 ** 16, 2 ** 20, 1000003]
meta['planner']['fourier_orders'] = small
meta['planner']['fourier_numL'] = [int(fm.Fourier(o)._numL) for o in small]
circ_n = list(range(1, 200)) + [251, 509, 1000, 1021, 4096, 2 ** 20]
meta['planner']['circulant_n'] = circ_n
meta['planner']['circulant_inner'] = [int(fm.Circulant(np.ones(n)).content[0].numRows) for n in circ_n]
toep = [(4, 3), (4, 5), (4, 41), (8, 8), (100, 100), (512, 512), (41, 4), (7, 7), (2, 2), (33, 65), (1, 5), (5, 1)]
meta['planner']['toeplitz_nm'] = toep
meta['planner']['toeplitz_inner'] = [
    int(fm.Toeplitz(np.ones(n), np.ones(m - 1)).content[0].numRows) for n, m in toep]

# ------------------------------------------------------------------ Fourier
for n in (1, 2, 3, 4, 8, 16, 35, 64, 127, 128, 243, 256):
    for optimize in (True, False):
        if not optimize and n not in (35, 127):
            continue
        F = fm.Fourier(n, optimize=optimize)
        x = crand(n, 3)
        put('fourier_n%d_opt%d' % (n, optimize), 'fourier', {'n': n, 'optimize': optimize},
            x=x, fwd=F.forward(x), bwd=F.backward(x))
for n, optimize in ((1000, True), (1021, True), (1021, False), (1024, True), (4096, True), (6144, True), (8192, True),
                    (2 ** 14, True), (2 ** 16, True), (2 ** 18, True), (2 ** 20, True), (3 * 2 ** 15, True),
                    (1000003, True), (5 ** 7, True)):
    x = seeded(n, n, 2)
    F = fm.Fourier(n, optimize=optimize)
    put_big('fourier_big_n%d_opt%d' % (n, optimize), 'fourier_big',
            {'n': n, 'seed': n, 'cols': 2, 'numL': int(F._numL), 'optimize': optimize},
            fwd=F.forward(x), bwd=F.backward(x))
x1 = crand(35)
put('fourier_1d_n35', 'fourier', {'n': 35, 'optimize': True}, x=x1, fwd=fm.Fourier(35).forward(x1),
    bwd=fm.Fourier(35).backward(x1))

# ------------------------------------------------------------------ Circulant (1 level; generator cast to double)
for n in (1, 2, 7, 31, 41, 64, 100, 127):
    c = crand(n)
    C = fm.Circulant(c)
    x = crand(n, 3)
    put('circulant_n%d' % n, 'circulant', {'n': n, 'inner': int(C.content[0].numRows)},
        c=c, x=x, fwd=C.forward(x), bwd=C.backward(x))
c = rng.integers(-3, 4, 41).astype(np.int32)
C = fm.Circulant(c.astype(np.float64))
x = rng.standard_normal((41, 2))
put('circulant_real_n41', 'circulant', {'n': 41, 'inner': int(C.content[0].numRows)},
    c=c.astype(np.float64), x=x, fwd=C.forward(x), bwd=C.backward(x))
c = crand(41)
C = fm.Circulant(c, optimize=False)
x = crand(41, 2)
put('circulant_n41_noopt', 'circulant', {'n': 41, 'inner': 41, 'optimize': False},
    c=c, x=x, fwd=C.forward(x), bwd=C.backward(x))
for n in (1000, 1021, 4096, 2 ** 17, 2 ** 20, 100003):
    c = seeded(n + 1, n)
    C = fm.Circulant(c)
    x = seeded(n, n, 2)
    put_big('circulant_big_n%d' % n, 'circulant_big',
            {'n': n, 'seed_c': n + 1, 'seed': n, 'cols': 2, 'inner': int(C.content[0].numRows)},
            fwd=C.forward(x), bwd=C.backward(x))

# ------------------------------------------------------------------ Circulant multi-level
for shape in ((3, 4), (3, 4, 5), (8, 16)):
    c = crand(*shape)
    C = fm.Circulant(c)
    n = int(np.prod(shape))
    x = crand(n, 2)
    put('circulant_ml_%s' % 'x'.join(map(str, shape)), 'circulant_ml', {'shape': list(shape)},
        c=c, x=x, fwd=C.forward(x), bwd=C.backward(x))

# ------------------------------------------------------------------ Toeplitz (1 level)
for n, m in ((4, 3), (4, 5), (4, 41), (8, 8), (41, 4), (100, 100), (1, 5), (5, 1)):
    vc = crand(n)
    vr = crand(m - 1)
    T = fm.Toeplitz(vc, vr)
    x = crand(m, 3)
    y = crand(n, 3)
    put('toeplitz_%dx%d' % (n, m), 'toeplitz', {'n': n, 'm': m, 'inner': int(T.content[0].numRows)},
        vc=vc, vr=vr, x=x, y=y, fwd=T.forward(x), bwd=T.backward(y))
for n, m in ((512, 512), (1000, 24), (2048, 2048), (2 ** 16, 2 ** 16), (2 ** 19, 2 ** 19), (70000, 50000)):
    vc = seeded(n + 1, n)
    vr = seeded(m + 2, m - 1)
    T = fm.Toeplitz(vc, vr)
    x = seeded(m, m, 2)
    y = seeded(n + 7, n, 2)
    put_big('toeplitz_big_%dx%d' % (n, m), 'toeplitz_big',
            {'n': n, 'm': m, 'seed_c': n + 1, 'seed_r': m + 2, 'seed_x': m, 'seed_y': n + 7, 'cols': 2,
             'inner': int(T.content[0].numRows)}, fwd=T.forward(x), bwd=T.backward(y))

# ------------------------------------------------------------------ Toeplitz multi-level
tt = crand(3, 3, 41)
T = fm.Toeplitz(tt, split=[2, 1, 4])
x = crand(T.numCols, 2)
y = crand(T.numRows, 2)
put('toeplitz_ml_3x3x41', 'toeplitz_ml', {'shape': [3, 3, 41], 'split': [2, 1, 4]},
    t=tt, x=x, y=y, fwd=T.forward(x), bwd=T.backward(y))
tt = crand(5, 7)
T = fm.Toeplitz(tt)
x = crand(T.numCols, 2)
y = crand(T.numRows, 2)
put('toeplitz_ml_5x7', 'toeplitz_ml', {'shape': [5, 7], 'split': None},
    t=tt, x=x, y=y, fwd=T.forward(x), bwd=T.backward(y))

# ------------------------------------------------------------------ Hadamard: all 8 dtypes, as-is (bit-exact contract)
for order in (1, 2, 4, 6, 10):
    H = fm.Hadamard(order)
    n = 2 ** order
    for dt in ALL_TYPES:
        if dt.startswith('int'):
            x = rng.choice(np.array([-2, -1, 1, 2]), size=(n, 3)).astype(dt)
        elif dt.startswith('float'):
            x = rng.standard_normal((n, 3)).astype(dt)
        else:
            x = crand(n, 3).astype(dt)
        put('hadamard_o%d_%s' % (order, dt), 'hadamard', {'order': order, 'dtype': dt}, x=x, fwd=H.forward(x))
x = np.full((1024, 1), 100, dtype=np.int8)                      # wrap-around KAT
put('hadamard_wrap_int8', 'hadamard', {'order': 10, 'dtype': 'int8'}, x=x, fwd=fm.Hadamard(10).forward(x))
def seeded_typed(seed, dt, *shape):
    """Seeded full-range input of any of the 8 fastmat dtypes (tests regenerate it with this exact recipe)."""
    g = np.random.default_rng(seed)
    if dt.startswith('int'):
        info = np.iinfo(dt)
        return g.integers(info.min, info.max, size=shape, dtype=np.int64, endpoint=True).astype(dt)
    if dt.startswith('float'):
        return g.standard_normal(shape).astype(dt)
    return (g.standard_normal(shape) + 1j * g.standard_normal(shape)).astype(dt)


for order, dts in ((13, ALL_TYPES), (16, ALL_TYPES), (20, ['int32', 'float32', 'int8', 'complex64'])):
    n = 2 ** order
    for dt in dts:
        x = seeded_typed(order, dt, n, 2)
        put_big('hadamard_big_o%d_%s' % (order, dt), 'hadamard_big',
                {'order': order, 'dtype': dt, 'seed': order, 'cols': 2}, fwd=fm.Hadamard(order).forward(x))

# ------------------------------------------------------------------ Kron of Fouriers / mixed
x = seeded(1024, 2 ** 20, 1)
K = fm.Kron(fm.Fourier(1024), fm.Fourier(1024))
put_big('kronF_big_1024x1024', 'kron_fourier_big', {'dims': [1024, 1024], 'seed': 1024, 'cols': 1},
        fwd=K.forward(x), bwd=K.backward(x))
for dims in ((4, 8), (3, 5, 4), (16, 16), (8, 32), (5, 7)):
    K = fm.Kron(*[fm.Fourier(d) for d in dims])
    n = int(np.prod(dims))
    x = crand(n, 2)
    put('kronF_%s' % 'x'.join(map(str, dims)), 'kron_fourier', {'dims': list(dims)},
        x=x, fwd=K.forward(x), bwd=K.backward(x))
K = fm.Kron(fm.Hadamard(3), fm.Fourier(5))
x = crand(40, 2)
put('kron_H3_F5', 'kron_hf', {'order': 3, 'n': 5}, x=x, fwd=K.forward(x), bwd=K.backward(x))
A = [rng.standard_normal((5, 5)), rng.standard_normal((4, 4)), rng.standard_normal((3, 3))]
K = fm.Kron(*[fm.Matrix(a) for a in A])
x = rng.standard_normal((60, 2))
put('kron_dense_5x4x3', 'kron_dense', {}, a0=A[0], a1=A[1], a2=A[2], x=x, fwd=K.forward(x), bwd=K.backward(x))

# ------------------------------------------------------------------ Partial / Diag / Product / Permutation / Sum / Blocks
H = fm.Hadamard(4)
rows = np.array([1, 2, 3, 11, 12])
cols = np.array([3, 7, 9, 15])
P = fm.Partial(H, rows=rows, cols=cols)
x = rng.integers(-5, 6, size=(4, 3)).astype(np.int32)
y = rng.integers(-5, 6, size=(5, 3)).astype(np.int32)
put('partial_hadamard4', 'partial_hadamard', {'order': 4}, rows=rows, cols=cols, x=x, y=y,
    fwd=P.forward(x), bwd=P.backward(y))
mask = rng.random(16) < 0.5
P = fm.Partial(H, rows=mask)
x = rng.standard_normal((16, 2))
y = rng.standard_normal((int(mask.sum()), 2))
put('partial_hadamard4_bool', 'partial_hadamard_bool', {'order': 4}, rows=mask, x=x, y=y,
    fwd=P.forward(x), bwd=P.backward(y))

n = 256
idx = np.sort(rng.choice(n, size=64, replace=False))
d = np.exp(2j * np.pi * rng.random(n))
A = fm.Product(fm.Partial(fm.Fourier(n), rows=idx), fm.Diag(d))
x = crand(n, 3)
y = crand(64, 3)
put('cs_partial_fourier_diag', 'cs_operator', {'n': n}, idx=idx, d=d, x=x, y=y, fwd=A.forward(x), bwd=A.backward(y))

for dt in ALL_TYPES:
    n = 35
    if dt.startswith('int'):
        d = rng.choice(np.array([-2, -1, 1, 2]), size=n).astype(dt)
    elif dt.startswith('float'):
        d = rng.standard_normal(n).astype(dt)
    else:
        d = crand(n).astype(dt)
    D = fm.Diag(d)
    for dx in ALL_TYPES:
        if dx.startswith('int'):
            x = rng.choice(np.array([-2, -1, 1, 2]), size=(n, 2)).astype(dx)
        elif dx.startswith('float'):
            x = rng.standard_normal((n, 2)).astype(dx)
        else:
            x = crand(n, 2).astype(dx)
        put('diag_%s_%s' % (dt, dx), 'diag', {'dtype_d': dt, 'dtype_x': dx}, d=d, x=x,
            fwd=D.forward(x), bwd=D.backward(x))

sigma = rng.permutation(35)
Pm = fm.Permutation(sigma)
x = rng.integers(-100, 100, size=(35, 3)).astype(np.int64)
put('permutation_35', 'permutation', {}, sigma=sigma, x=x, fwd=Pm.forward(x), bwd=Pm.backward(x))

c = crand(16)
d = crand(16)
S = fm.Sum(fm.Circulant(c), fm.Diag(d), fm.Fourier(16))
x = crand(16, 2)
put('sum_circ_diag_fourier', 'sum', {}, c=c, d=d, x=x, fwd=S.forward(x), bwd=S.backward(x))
B = fm.Blocks([[fm.Circulant(c), fm.Fourier(16)], [fm.Diag(d), fm.Hadamard(4)]])
x = crand(32, 2)
put('blocks_2x2', 'blocks', {}, c=c, d=d, x=x, fwd=B.forward(x), bwd=B.backward(x))
BD = fm.BlockDiag(fm.Fourier(16), fm.Hadamard(4))
put('blockdiag', 'blockdiag', {}, x=x, fwd=BD.forward(x), bwd=BD.backward(x))
Pr = fm.Product(fm.Hadamard(4), 2.5 - 1j, fm.Diag(d), fm.Fourier(16))
x = crand(16, 2)
put('product_scalar', 'product', {'scalar': [2.5, -1.0]}, d=d, x=x, fwd=Pr.forward(x), bwd=Pr.backward(x))
Fh = fm.Fourier(16)
put('fourier_views', 'views', {}, x=x, H=Fh.H.forward(x), T=Fh.T.forward(x), conj=Fh.conj.forward(x))

# ------------------------------------------------------------------ as-is dtype behaviour of the reference (this numpy)
for name, M, n in (('Fourier', fm.Fourier(16), 16), ('Hadamard', fm.Hadamard(4), 16),
                   ('Circulant_c64', fm.Circulant(crand(16).astype(np.complex64)), 16),
                   ('Circulant_f64', fm.Circulant(rng.standard_normal(16)), 16),
                   ('Toeplitz_c64', fm.Toeplitz(crand(16).astype(np.complex64), crand(15).astype(np.complex64)), 16),
                   ('Diag_f32', fm.Diag(rng.standard_normal(16).astype(np.float32)), 16),
                   ('KronFF', fm.Kron(fm.Fourier(4), fm.Fourier(4)), 16)):
    meta['dtypes'][name] = {dt: str(M.forward(np.ones((n, 1), dtype=dt)).dtype) for dt in ALL_TYPES}

np.savez_compressed(os.path.join(OUT, 'golden.npz'), **arrays)
with open(os.path.join(OUT, 'golden.json'), 'w') as f:
    json.dump(meta, f, indent=1, sort_keys=True)
print('wrote %d arrays, %d cases, %.2f MB' % (len(arrays), len(meta['cases']),
                                               os.path.getsize(os.path.join(OUT, 'golden.npz')) / 1e6))
