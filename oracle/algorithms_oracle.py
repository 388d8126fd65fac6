"""CPU restatement (numpy, dense matrices) of fastmat's sparse-recovery solvers -- TEST INFRASTRUCTURE ONLY.

Nothing under fastmat_b200/ imports this file.  Each function follows the reference statement by statement and cites
it; the system matrix is a dense ndarray (forward = A @ x, backward = A^H @ y), which is what the reference's own
``Matrix`` class does for dense input (fastmat/Matrix.pyx:1831-1842, :1924-1935).  Pinned against the real reference
by tests/test_algorithms_cpu.py through tests/golden/golden_algorithms.npz (made by oracle/make_golden_algorithms.py).
"""
import numpy as np


def soft_threshold(x, alpha):
    """fastmat/algorithms/ISTA.py:113-123."""
    m = np.maximum(np.abs(x) - alpha, 0)
    return np.multiply(m / (m + alpha), x)


def largest_singular_value(A):
    """fastmat/Matrix.pyx:895-919 (scipy svds, k=1) == spectral norm."""
    return float(np.linalg.norm(A, 2))


def ista(A, b, numLambda=0.1, numMaxSteps=100, lsv=None):
    """fastmat/algorithms/ISTA.py:125-165."""
    b2 = b.reshape(-1, 1) if b.ndim == 1 else b
    numL = 1.0 / ((largest_singular_value(A) if lsv is None else lsv) ** 2)
    x = np.zeros((A.shape[1], b2.shape[1]), dtype=np.promote_types(np.float32, b2.dtype))
    step = x
    for _ in range(numMaxSteps):
        step = x - numL * (A.conj().T @ (A @ x - b2))
        x = soft_threshold(step, numL * numLambda * 0.5)
    return np.where(x != 0, step, x)


def fista(A, b, numLambda=0.1, numMaxSteps=100, lsv=None):
    """fastmat/algorithms/FISTA.py:131-170."""
    b2 = b.reshape(-1, 1) if b.ndim == 1 else b
    numL = 1.0 / ((largest_singular_value(A) if lsv is None else lsv) ** 2)
    t = 1
    x = np.zeros((A.shape[1], b2.shape[1]), dtype=np.promote_types(np.float32, b2.dtype))
    y = np.copy(x)
    step = x
    for _ in range(numMaxSteps):
        xold = np.copy(x)
        step = y - numL * (A.conj().T @ (A @ y - b2))
        x = soft_threshold(step, numL * numLambda * 0.5)
        told = t
        t = (1 + np.sqrt(1 + 4 * t ** 2)) / 2
        y = x + ((told - 1) / t) * (x - xold)
    return np.where(x != 0, step, x)


def omp(A, b, numMaxSteps):
    """fastmat/algorithms/OMP.pyx:122-253."""
    b2 = b.reshape(-1, 1) if b.ndim == 1 else b
    N, M = A.shape
    L = b2.shape[1]
    K = numMaxSteps
    C = A / np.linalg.norm(A, axis=0)                                   # colNormalized, fastmat/Matrix.pyx:1090-1120
    rt = np.promote_types(np.promote_types(C.dtype, b2.dtype), np.float64)
    xtmp = np.zeros((K, L), dtype=rt)
    res = b2.astype(rt, copy=True)
    support = np.empty((K, L), dtype=np.intp)
    pinv = np.zeros((K, N, L), dtype=rt)
    arrA = np.zeros((N, K, L), dtype=rt)
    for ii in range(K):
        c = np.abs(C.conj().T @ res)
        idx = np.apply_along_axis(np.argmax, 0, c)
        support[ii, :] = idx
        newcols = A[:, idx]
        arrA[:, ii, :] = newcols
        if ii == 0:
            v2 = newcols
            v2n = (v2 / np.linalg.norm(v2, axis=0) ** 2).conj()
            v2y = np.einsum('ji,ji->i', v2n, b2)
            xtmp[0, :] = v2y
            pinv[0, :, :] = v2n
        else:
            v1 = np.einsum('ijk,jk->ik', pinv[:ii, :, :], newcols)
            v2 = newcols - np.einsum('ijk,jk->ik', arrA[:, :ii, :], v1)
            v2n = (v2 / np.linalg.norm(v2, axis=0) ** 2).conj()
            v2y = np.einsum('ji,ji->i', v2n, b2)
            xtmp[:ii, :] -= v2y * v1
            xtmp[ii, :] += v2y
            pinv[:ii, :, :] -= np.einsum('ik,jk->jik', v2n, v1)
            pinv[ii, :, :] = v2n
        res -= v2y * v2
    x = np.zeros((M, L), dtype=rt)
    x[support, np.arange(L)] = xtmp
    return x


def stela(A, b, numLambda=0.1, numMaxSteps=100, numMaxError=1e-6):
    """fastmat/algorithms/STELA.py:128-262."""
    b2 = b.reshape(-1, 1) if b.ndim == 1 else b
    dt = np.promote_types(np.float64, b2.dtype)
    gamma = np.zeros(b2.shape[1], dtype=dt)
    x = np.zeros((A.shape[1], b2.shape[1]), dtype=dt)
    res = (-b2).astype(dt)
    bx = np.zeros_like(x, dtype=dt)
    abxx = np.zeros_like(b2, dtype=dt)
    AH = A.conj().T
    z = AH @ res
    d = (1.0 / np.linalg.norm(A, axis=0) ** 2).reshape((-1, 1))
    for _ in range(numMaxSteps):
        grad = d * x - z
        diff = np.maximum(np.minimum(z.real - x.real, +numLambda), -numLambda)
        if dt == complex:
            diff = diff + 1j * np.maximum(np.minimum(z.imag - x.imag, +numLambda), -numLambda)
        stop = np.linalg.norm(z - diff, axis=0)
        act = stop > numMaxError
        if np.sum(act) == 0:
            return x
        bx[:, act] = soft_threshold(grad[:, act], numLambda) / d
        abxx[:, act] = A @ (bx[:, act] - x[:, act])
        gamma[act] = np.maximum(np.minimum(
            -(np.real(np.sum(np.multiply(np.conj(res[:, act]), abxx[:, act]), axis=0))
              + numLambda * (np.sum(np.abs(bx[:, act]) - np.abs(x[:, act]), axis=0)))
            / np.sum(np.abs(abxx[:, act]) ** 2, axis=0), 1), 0)
        x[:, act] += (bx[:, act] - x[:, act]).dot(np.diag(gamma[act]))
        res[:, act] += gamma[act] * abxx[:, act]
        z[:, act] = AH @ res[:, act]
    return x


# ---- the compressed-sensing operator of BASELINE config 5, as a dense matrix
def cs_matrix_fourier(n, rows, d):
    """dense Product(Partial(Fourier(n), rows=rows), Diag(d)) (fastmat/Fourier.pyx:241-246, Partial.pyx:296-307, Diag.pyx:169-176)."""
    k = np.arange(n)
    F = np.exp(-2j * np.pi * np.outer(k, k) / n)
    return F[rows, :] * d[None, :]


def cs_matrix_hadamard(order, rows, d):
    """dense Product(Partial(Hadamard(order), rows=rows), Diag(d)) (fastmat/Hadamard.pyx:242-248)."""
    H = np.array([[1.0]])
    for _ in range(order):
        H = np.block([[H, H], [H, -H]])
    return H[rows, :] * d[None, :]
