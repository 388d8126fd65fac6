"""CPU: the numpy restatement of ISTA / FISTA / OMP (oracle/algorithms_oracle.py) against fixtures frozen from the real
reference's solvers (tests/golden/golden_algorithms.npz <- oracle/make_golden_algorithms.py), plus the host-side logic
of the solver classes that needs no GPU (parameter handling, error behaviour)."""
import os

import numpy as np
import pytest

from oracle import algorithms_oracle as ao

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GA = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_algorithms.npz'))


def problem(tag):
    if tag == 'cs':
        A = ao.cs_matrix_fourier(256, GA['cs_rows'], GA['cs_d'])
    else:
        A = ao.cs_matrix_hadamard(8, GA['had_rows'], GA['had_d'])
    lam, steps, k = GA[tag + '_params']
    return A, GA[tag + '_b'], float(lam), int(steps), int(k)


@pytest.mark.parametrize('tag', ['cs', 'had'])
def test_oracle_operator_matches_reference(tag):
    A, b, _, _, _ = problem(tag)
    assert np.abs(A @ GA[tag + '_x'] - b).max() <= 1e-10 * np.abs(b).max()
    assert abs(ao.largest_singular_value(A) - float(GA[tag + '_lsv'])) <= 1e-8 * float(GA[tag + '_lsv'])


@pytest.mark.parametrize('tag', ['cs', 'had'])
@pytest.mark.parametrize('alg', ['ista', 'fista'])
def test_oracle_ista_matches_reference(tag, alg):
    A, b, lam, steps, _ = problem(tag)
    got = getattr(ao, alg)(A, b, numLambda=lam, numMaxSteps=steps, lsv=float(GA[tag + '_lsv']))
    ref = GA['%s_%s' % (tag, alg)]
    assert got.shape == ref.shape
    assert np.array_equal(got != 0, ref != 0)                       # same support after thresholding
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()


@pytest.mark.parametrize('tag', ['cs', 'had'])
def test_oracle_omp_matches_reference(tag):
    A, b, _, _, k = problem(tag)
    got = ao.omp(A, b, k)
    ref = GA[tag + '_omp']
    assert np.array_equal(got != 0, ref != 0)
    assert np.abs(got - ref).max() <= 1e-10 * np.abs(ref).max()
    # and OMP recovers the k-sparse ground truth exactly (the reference's own criterion, fastmat/test/algorithm.py:58-64)
    x = GA[tag + '_x']
    assert np.array_equal(got != 0, x != 0)
    assert np.allclose(got, x, rtol=1e-9, atol=1e-9)


def test_soft_threshold_definition():
    x = np.array([-3.0, -0.5, 0.0, 0.2, 2.0])
    y = ao.soft_threshold(x, 1.0)
    assert np.allclose(y, [-2.0, 0.0, 0.0, 0.0, 1.0])
    z = ao.soft_threshold(np.array([3 + 4j]), 1.0)                  # shrinks the modulus, keeps the phase
    assert np.allclose(z, (3 + 4j) * 4.0 / 5.0)


@pytest.mark.parametrize('tag', ['cs', 'had'])
def test_oracle_stela_matches_reference(tag):
    A, b, lam, _, _ = problem(tag)
    got = ao.stela(A, b, numLambda=lam, numMaxSteps=8)
    ref = GA[tag + '_stela']
    assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
