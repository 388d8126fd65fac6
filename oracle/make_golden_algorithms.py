#!/usr/bin/env python
"""Generate tests/golden/golden_algorithms.npz from the REAL reference's solvers (TEST INFRASTRUCTURE ONLY).

    bash oracle/build_ref.sh && python oracle/make_golden_algorithms.py

The reference's own solver tests are weak (ISTA/FISTA: CHECK_PROXIMITY off; OMP: exact support on a dense Gaussian
matrix, fastmat/test/algorithm.py:58-64), so the fixtures are seeded runs of fastmat.algorithms.{ISTA,FISTA,OMP} on small
instances of BASELINE config 5 (Product(Partial(Fourier), Diag)) and of a real-valued Hadamard variant, in double.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '_ref'))
import fastmat as fm                                    # noqa: E402  (the real reference)
import fastmat.algorithms as fma                        # noqa: E402

OUT = os.path.join(HERE, '..', 'tests', 'golden', 'golden_algorithms.npz')
rng = np.random.default_rng(777)
arrs = {}


def sparse_truth(n, L, k, cplx):
    x = np.zeros((n, L), dtype=complex if cplx else float)
    for c in range(L):
        idx = rng.choice(n, k, replace=False)
        x[idx, c] = (rng.standard_normal(k) + (1j * rng.standard_normal(k) if cplx else 0)) + 2 * np.sign(rng.standard_normal(k))
    return x


def run(tag, A, x, lam, steps, k):
    b = A.forward(x)
    arrs[tag + '_x'] = x
    arrs[tag + '_b'] = b
    arrs[tag + '_lsv'] = np.asarray(A.largestSingularValue)
    arrs[tag + '_ista'] = fma.ISTA(A, numLambda=lam, numMaxSteps=steps).process(b)
    arrs[tag + '_fista'] = fma.FISTA(A, numLambda=lam, numMaxSteps=steps).process(b)
    arrs[tag + '_omp'] = fma.OMP(A, numMaxSteps=k).process(b)
    # STELA amplifies rounding differences step by step (its line-search steps are tiny on these problems): 8 steps
    arrs[tag + '_stela'] = fma.STELA(A, numLambda=lam, numMaxSteps=8).process(b)
    arrs[tag + '_params'] = np.asarray([lam, steps, k], dtype=float)


# config-5 shape, small: n = 256, 64 of the Fourier rows, unit-modulus diagonal, 4-sparse, 6 right-hand sides
n = 256
rows = np.sort(rng.choice(n, 64, replace=False))
d = np.exp(2j * np.pi * rng.random(n))
arrs['cs_rows'] = rows
arrs['cs_d'] = d
A = fm.Product(fm.Partial(fm.Fourier(n), rows=rows), fm.Diag(d))
run('cs', A, sparse_truth(n, 6, 4, True), 0.05, 60, 4)

# real-valued: Hadamard(8), 96 rows, positive diagonal
order = 8
rows = np.sort(rng.choice(1 << order, 96, replace=False))
d = 0.5 + rng.random(1 << order)
arrs['had_rows'] = rows
arrs['had_d'] = d
A = fm.Product(fm.Partial(fm.Hadamard(order), rows=rows), fm.Diag(d))
run('had', A, sparse_truth(1 << order, 5, 3, False), 0.1, 40, 3)

np.savez_compressed(OUT, **arrs)
print('wrote', OUT, {k: v.shape for k, v in arrs.items()})
