"""CPU oracle for the fastmat structured-matrix apply path -- TEST INFRASTRUCTURE ONLY.

This module is a plain-numpy restatement of the reference algorithms (EMS-TU-Ilmenau/fastmat v0.2.2.post0,
read-only at /root/reference).  It exists so that ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline
legs of ``bench.py`` have something that travels to the GPU box (the reference itself does not).  Nothing under
``fastmat_b200/`` may import it: the product path is CUDA-only and fails loudly without its extension.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the real reference (built by ``oracle/build_ref.sh``
into ``oracle/_ref``) on seeded inputs and stores the outputs under ``tests/golden/``; ``tests/test_oracle.py``
checks every function below against those fixtures (planner tables bit-exact, integer Hadamard bit-exact,
floating point to ~1e-13 in double).

Third-party arithmetic: every FFT butterfly of the reference is numpy's bundled pocketfft (``np.fft.fft``;
numpy is unpinned in the reference, 2.3.5 here).  The oracle calls the same ``np.fft`` -- by design: the oracle
restates fastmat, not pocketfft.  Independent ground truth for the transforms is the dense ``reference()``
construction (``dense_*`` below), exactly as the reference's own tests do (fastmat/inspect/test.py:318-368).

Precision policy: the reference up-casts to complex128 inside Product (fastmat/Product.pyx:215); numpy>=2 keeps
single precision inside np.fft for complex64 inputs.  All oracle functions therefore take ``double=True`` by
default = "evaluate in complex128/float64" (ground truth for the tolerance tests, SURVEY.md section 8c rule 1).

Each function cites the reference file:line it follows.
"""
import numpy as np

# ----------------------------------------------------------------------------------------------- planner

def find_fft_factors(target_length, max_factor, state, best_state):
    """fastmat/core/cmath.pyx:35-85 (_findFFTFactors): greedy recursion on (complexity<<16)+length states."""
    for ff in range(max_factor, 0, -1):
        length = (state & 0xFFFF) * ff
        complexity = (state >> 16) + ff + 1
        new_state = (complexity << 16) + length
        if new_state <= best_state and length < target_length:
            best_state = find_fft_factors(target_length, ff, new_state, best_state)
        else:
            if new_state < best_state:
                best_state = new_state
    return best_state


def find_optimal_fft_size(order, max_stage=4):
    """fastmat/core/cmath.pyx:88-153 (_findOptimalFFTSize).

    ``remaining`` is a C ``float`` in the reference (:131); np.float32 reproduces the rounding of the
    int -> float conversion (division by 4 is exact), including the wrong answer above 2**24.
    """
    padded = 1
    remaining = np.float32(order)
    while remaining > np.float32(64):
        padded *= 4
        remaining = np.float32(remaining / np.float32(4))
    x = int(np.ceil(remaining))
    if x != 1:
        length = 64
        complexity = 3 * (4 + 1)
        factor = find_fft_factors(x, max_stage, 1, (complexity << 16) + length) & 0xFFFF
        padded *= factor
    return int(padded)


def get_fft_complexity(n):
    """fastmat/core/cmath.pyx:156-214 (_getFFTComplexity); float32 accumulator and result (:181, :214)."""
    n = int(n)
    complexity = np.float32(0)
    nn = n
    while nn % 4 == 0:
        complexity = np.float32(complexity + np.float32(5))
        nn //= 4
    if nn > 1 and (nn & 1) == 0:
        complexity = np.float32(complexity + np.float32(3))
        nn //= 2
    ii = 3
    while nn > 1 and ii * ii < nn:          # strict '<' (:202): a leftover p*p is charged as one factor
        if nn % ii == 0:
            complexity = np.float32(complexity + np.float32(ii + 1))
            nn //= ii
        else:
            ii += 2
    if nn > 1:
        complexity = np.float32(complexity + np.float32(nn + 1))
    return np.float32(np.float32(n) * np.float32(complexity + np.float32(1)))


def fourier_bluestein_size(order, optimize=True, max_stage=4):
    """fastmat/Fourier.pyx:109-122: 0 = plain FFT, else the chirp-z inner length L (``_numL``)."""
    if not optimize:
        return 0
    padded = find_optimal_fft_size(order * 2 - 1, max_stage)
    lhs = get_fft_complexity(order)
    # python arithmetic on C floats promotes to double in Cython's generated code for the mixed expression
    rhs = 2 * float(get_fft_complexity(padded)) + 2 * padded + 2 * order
    return 0 if float(lhs) < rhs else padded


def circulant_inner_size(n, optimize=True, max_stage=4):
    """fastmat/Circulant.pyx:104-124: size of the circulant the n x n matrix is embedded in."""
    if not optimize:
        return n
    padded = find_optimal_fft_size(2 * n - 1, max_stage)
    return padded if get_fft_complexity(n) > get_fft_complexity(padded) else n


def toeplitz_inner_size(n, m, optimize=True, max_stage=4):
    """fastmat/Toeplitz.pyx:225-233: FFT length for an n x m Toeplitz (defining vector length n+m-1)."""
    d = n + m - 1
    if not optimize:
        return d
    opt = find_optimal_fft_size(d, max_stage)
    return opt if get_fft_complexity(opt) < get_fft_complexity(d) else d


# ----------------------------------------------------------------------------------------------- types

FTYPES = [np.int8, np.int16, np.int32, np.int64, np.float32, np.float64, np.complex64, np.complex128]


def promote(a, b):
    """fastmat/core/types.pyx:443-453: the promotion table is np.promote_types over the 8 fastmat types."""
    return np.promote_types(a, b).type


def safe_type_expansion(dtype):
    """fastmat/core/types.pyx:378-394."""
    dtype = np.dtype(dtype).type
    if dtype in (np.int8, np.int16):
        return np.float32
    if dtype in (np.int32, np.int64):
        return np.float64
    return dtype


def _as2d(x):
    x = np.asarray(x)
    if x.ndim == 1:
        return x.reshape(-1, 1), True
    if x.ndim != 2:
        raise ValueError("Input data array must be 1D or 2D")      # fastmat/Matrix.pyx:1772
    return x, False


def _ret(y, was1d):
    return y.reshape(-1) if was1d else y


def _cplx(x, double):
    if double:
        return np.asarray(x).astype(np.complex128)
    return np.asarray(x)


# ----------------------------------------------------------------------------------------------- Fourier

def fourier_forward(x, optimize=True, max_stage=4, double=True):
    """fastmat/Fourier.pyx:208-233.  Plain np.fft.fft along axis 0, or chirp-z when ``_numL`` > 0."""
    x2, was1d = _as2d(x)
    n = x2.shape[0]
    xin = _cplx(x2, double)
    L = fourier_bluestein_size(n, optimize, max_stage)
    if L == 0:
        y = np.fft.fft(xin, axis=0)                                         # :214
    else:
        # __init__ :127-156
        k = np.linspace(0, n - 1, n)
        arg = (k ** 2) * np.pi / n
        pre = np.cos(arg) - 1j * np.sin(arg)
        conv = np.zeros(L, dtype=np.complex128)
        conv[:n] = np.cos(arg) + 1j * np.sin(arg)
        conv[L - n + 1:] = conv[1:n][::-1]
        conv_hat = np.fft.fft(conv)
        # _forward :215-231
        buf = np.zeros((L, x2.shape[1]), dtype=np.promote_types(np.complex128, xin.dtype))
        buf[:n] = pre[:, None] * xin
        buf = conv_hat[:, None] * np.fft.fft(buf, axis=0)
        y = pre[:, None] * np.fft.ifft(buf, axis=0)[:n]
    return _ret(y, was1d)


def fourier_backward(x, optimize=True, max_stage=4, double=True):
    """fastmat/Fourier.pyx:235-238: conj(F conj(x))."""
    x2, was1d = _as2d(x)
    y = np.conj(fourier_forward(np.conj(_cplx(x2, double)), optimize, max_stage, double))
    return _ret(y, was1d)


def dense_fourier(n):
    """fastmat/Fourier.pyx:241-246 (_reference): DFT matrix from np.exp, independent of np.fft."""
    k = np.arange(n)
    return np.exp(np.multiply(*np.meshgrid(k, k)) * (-2j * np.pi / n))


# ----------------------------------------------------------------------------------------------- Diag

def diag_forward(d, x):
    """fastmat/Diag.pyx:149-157 -> core/cmath.pyx:989-997: out[n, m] = x[n, m] * d[n], promoted dtype."""
    x2, was1d = _as2d(x)
    d = np.asarray(d)
    out = (x2.astype(np.promote_types(x2.dtype, d.dtype)) * d[:, None]).astype(np.promote_types(x2.dtype, d.dtype))
    return _ret(out, was1d)


def diag_backward(d, x):
    """fastmat/Diag.pyx:159-167: multiply by conj(d)."""
    return diag_forward(np.conj(np.asarray(d)), x)


# ----------------------------------------------------------------------------------------------- Circulant

def circulant_spectrum(c, optimize=True, max_stage=4, double=True):
    """fastmat/Circulant.pyx:104-131: (inner size, fft(embedded c)/size)."""
    c = np.asarray(c).reshape(-1)
    n = c.size
    size = circulant_inner_size(n, optimize, max_stage)
    if size != n:
        c = np.concatenate([c, np.zeros(size - (2 * n - 1), dtype=c.dtype), c[1:]])     # :112-124
    cc = c.astype(np.complex128) if double else c
    return size, np.fft.fft(cc, axis=0) / size                                          # :131


def _circ_apply(spec, size, n_in, n_out, x2, double, adjoint):
    xin = _cplx(x2, double)
    buf = np.zeros((size, x2.shape[1]), dtype=np.promote_types(xin.dtype, np.complex64))
    buf[:n_in] = xin                                                   # Partial scatter, Partial.pyx:272-274
    s = np.conj(spec) if adjoint else spec
    # Product(F.H, D, F): F unnormalised, F.H = conj(F conj(.)) unnormalised, 1/size sits in D (:126-133)
    y = np.fft.fft(buf, axis=0) * s[:, None]
    y = np.conj(np.fft.fft(np.conj(y), axis=0))
    return y[:n_out]                                                   # Partial gather, Partial.pyx:277-280


def circulant_forward(c, x, optimize=True, max_stage=4, double=True):
    """fastmat/Circulant.pyx:63-227 graph Partial(Product(F.H, Diag, F)) applied forward."""
    x2, was1d = _as2d(x)
    n = np.asarray(c).size
    size, spec = circulant_spectrum(c, optimize, max_stage, double)
    return _ret(_circ_apply(spec, size, n, n, x2, double, False), was1d)


def circulant_backward(c, x, optimize=True, max_stage=4, double=True):
    """Product._backward (fastmat/Product.pyx:223-240): F.H.backward=F.forward, D.backward=conj(d), F.backward."""
    x2, was1d = _as2d(x)
    n = np.asarray(c).size
    size, spec = circulant_spectrum(c, optimize, max_stage, double)
    return _ret(_circ_apply(spec, size, n, n, x2, double, True), was1d)


def dense_circulant(c):
    """fastmat/Circulant.pyx:349-425 (_refRecursion, one level): C[i, j] = c[(i - j) mod n]."""
    c = np.asarray(c).reshape(-1)
    n = c.size
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    return c[(i - j) % n]


# ----------------------------------------------------------------------------------------------- Toeplitz

def toeplitz_spectrum(vec_c, vec_r, optimize=True, max_stage=4):
    """fastmat/Toeplitz.pyx:162-171, 225-260, 331-369: (L, fft(t_pad)/L); t = [vecC, vecR], zeros at the split."""
    vec_c = np.asarray(vec_c).reshape(-1)
    vec_r = np.asarray(vec_r).reshape(-1)
    n, m = vec_c.size, vec_r.size + 1
    t = np.hstack((vec_c, vec_r)).astype(np.complex128)                # :240 astype('complex')
    L = toeplitz_inner_size(n, m, optimize, max_stage)
    if L > t.size:
        t = np.concatenate((t[:n], np.zeros(L - t.size, dtype=t.dtype), t[n:]))         # _preProcSlice
    return L, np.fft.fft(t, axis=0) / L


def toeplitz_forward(vec_c, vec_r, x, optimize=True, max_stage=4, double=True):
    """fastmat/Toeplitz.pyx:79-322: rows < n, cols < m of the L x L circulant (:284-316)."""
    x2, was1d = _as2d(x)
    n, m = np.asarray(vec_c).size, np.asarray(vec_r).size + 1
    L, spec = toeplitz_spectrum(vec_c, vec_r, optimize, max_stage)
    return _ret(_circ_apply(spec, L, m, n, x2, double, False), was1d)


def toeplitz_backward(vec_c, vec_r, x, optimize=True, max_stage=4, double=True):
    x2, was1d = _as2d(x)
    n, m = np.asarray(vec_c).size, np.asarray(vec_r).size + 1
    L, spec = toeplitz_spectrum(vec_c, vec_r, optimize, max_stage)
    return _ret(_circ_apply(spec, L, n, m, x2, double, True), was1d)


def dense_toeplitz(vec_c, vec_r):
    """fastmat/Toeplitz.pyx:726-732: T[i, j] = t[(i - j) mod (n + m - 1)], t = [vecC, vecR]."""
    vec_c = np.asarray(vec_c).reshape(-1)
    vec_r = np.asarray(vec_r).reshape(-1)
    n, m = vec_c.size, vec_r.size + 1
    t = np.hstack((vec_c, vec_r))
    i, j = np.meshgrid(np.arange(n), np.arange(m), indexing='ij')
    return t[(i - j) % (n + m - 1)]


# ----------------------------------------------------------------------------------------------- Hadamard

def hadamard_forward(x, order=None):
    """fastmat/Hadamard.pyx:164-230 + _hadamardCore :36-61.

    Natural-order (Sylvester) WHT, unnormalised; per column, ``order`` in-place radix-2 stages with butterfly
    distance 1, 2, 4, ...; arithmetic in the INPUT's own dtype (integers wrap; output dtype = promote(in, int8)
    = in).  The stage order matters for floating-point bit-exactness and is kept.
    """
    x2, was1d = _as2d(x)
    n = x2.shape[0]
    if order is None:
        order = int(np.log2(n))
    assert n == 1 << order
    y = np.array(x2, order='C', copy=True)
    m = y.shape[1]
    with np.errstate(over='ignore'):
        dist = 1
        for _ in range(order):
            # A = elements with bit `dist` clear, B = the partner with the bit set (strideSubgrid, :199-208)
            v = y.reshape(n // (2 * dist), 2, dist, m)          # a view: y is C-contiguous
            a = v[:, 0].copy()
            b = v[:, 1].copy()
            v[:, 0] = a + b                                      # array arithmetic stays in y.dtype (ints wrap)
            v[:, 1] = a - b
            dist <<= 1
    return _ret(y, was1d)


def dense_hadamard(order, dtype=np.int8):
    """fastmat/Hadamard.pyx:242-248: scipy.linalg.hadamard (Sylvester construction), rebuilt here without scipy."""
    h = np.array([[1]], dtype=dtype)
    for _ in range(order):
        h = np.block([[h, h], [h, -h]]).astype(dtype)
    return h


# ----------------------------------------------------------------------------------------------- Partial

def partial_forward(apply_nested, nested_cols, rows, cols, x):
    """fastmat/Partial.pyx:268-280: zero buffer, scatter x into cols, nested forward, gather rows."""
    x2, was1d = _as2d(x)
    if cols is not None:
        buf = np.zeros((nested_cols, x2.shape[1]), dtype=x2.dtype)
        buf[np.asarray(cols)] = x2
    else:
        buf = x2
    y = apply_nested(buf)
    if rows is not None:
        y = y[np.asarray(rows)]
    return _ret(y, was1d)


def partial_backward(apply_nested_h, nested_rows, rows, cols, x):
    """fastmat/Partial.pyx:282-294."""
    x2, was1d = _as2d(x)
    if rows is not None:
        buf = np.zeros((nested_rows, x2.shape[1]), dtype=x2.dtype)
        buf[np.asarray(rows)] = x2
    else:
        buf = x2
    y = apply_nested_h(buf)
    if cols is not None:
        y = y[np.asarray(cols)]
    return _ret(y, was1d)


# ----------------------------------------------------------------------------------------------- Kron

def kron_forward(applies, dims, x, out_dtype=None):
    """fastmat/Kron.pyx:267-303: chain of mode-i products on the row-major index (i1, ..., ik).

    ``applies[i]`` maps a (dims[i], K) array to a (dims[i], K) array (term.forward / term.backward).
    The reference's reshape choreography (C-order to (head, .), F-order to (n_i, .)) is restated literally.
    """
    x2, was1d = _as2d(x)
    n_total = int(np.prod(dims))
    num_vecs = x2.shape[1]
    data = x2 if out_dtype is None else x2.astype(out_dtype)            # _widenInputDatatype, Kron.pyx:129
    head = 1
    for f, ni in zip(applies, dims):
        head *= ni
        data = np.reshape(data, (head, num_vecs * n_total // head), order='C')
        data = np.reshape(data, (ni, num_vecs * n_total // ni), order='F')
        data = np.reshape(f(data), (head, num_vecs * n_total // head), order='F')
    return _ret(np.reshape(data, (n_total, num_vecs), order='C'), was1d)


def kron_fourier_forward(dims, x, double=True):
    """Kron(Fourier(d1), ..., Fourier(dk)).forward == N-D FFT of the row-major reshaped column."""
    x2, was1d = _as2d(x)
    xin = _cplx(x2, double)
    t = xin.reshape(tuple(dims) + (x2.shape[1],))
    y = np.fft.fftn(t, axes=tuple(range(len(dims))))
    return _ret(y.reshape(x2.shape[0], x2.shape[1]), was1d)


def kron_fourier_backward(dims, x, double=True):
    x2, was1d = _as2d(x)
    y = np.conj(kron_fourier_forward(dims, np.conj(_cplx(x2, double)), double))
    return _ret(y, was1d)


# ----------------------------------------------------------------------------------------------- Product

def product_forward(applies, scalar, dtype, x):
    """fastmat/Product.pyx:205-221: astype(promote(x, dtype)) (or scalar inner), then factors right-to-left."""
    x2, was1d = _as2d(x)
    if scalar != 1:
        r = np.inner(x2, scalar)
    else:
        r = x2.astype(np.promote_types(x2.dtype, dtype))
    for f in reversed(applies):
        r = f(r)
    return _ret(r, was1d)


def permutation_forward(sigma, x):
    """fastmat/Permutation.pyx:121-125: x[sigma, :]."""
    x2, was1d = _as2d(x)
    return _ret(x2[np.asarray(sigma)], was1d)


def permutation_backward(sigma, x):
    x2, was1d = _as2d(x)
    tau = np.argsort(np.asarray(sigma))
    return _ret(x2[tau], was1d)


# ----------------------------------------------------------------------------------------------- LFSRCirculant

def lfsr_order(polynomial):
    """fastmat/LFSRCirculant.pyx:191-196: index of the highest set bit of the characteristic polynomial."""
    order, mask = 0, 1
    while (~mask & polynomial) > mask:
        mask <<= 1
        order += 1
    return order


def lfsr_gen_step(state, polynomial, mask):
    """fastmat/LFSRCirculant.pyx:31-40: Fibonacci step, feedback = parity(state & polynomial) enters at bit `order`."""
    if bin(state & polynomial).count('1') & 1:
        state |= mask
    return state >> 1


def lfsr_tap_step(state, polynomial, mask):
    """fastmat/LFSRCirculant.pyx:42-48: Galois step (multiply by x modulo the polynomial)."""
    state <<= 1
    if state & mask:
        state ^= (polynomial | mask)
    return state


def lfsr_period(polynomial, start):
    """fastmat/LFSRCirculant.pyx:196-222 including the constructor's three ValueErrors."""
    order = lfsr_order(polynomial)
    if order > 31 or order < 1:
        raise ValueError("Only polynomials of order 1 to 31 are supported.")
    mask = 1 << order
    start &= mask - 1
    if start == 0:
        raise ValueError("Initial state must be non-zero.")
    state, period = lfsr_gen_step(start, polynomial, mask), 1
    while state != start:
        state = lfsr_gen_step(state, polynomial, mask)
        period += 1
        if period >= mask or state == 0:
            raise ValueError("Register configuration produces invalid sequence.")
    return period


def lfsr_sequences(polynomial, start):
    """Generator states, tap states (from 1) and the +1/-1 output over one period (:277-314, :343-395)."""
    n = lfsr_period(polynomial, start)
    mask = 1 << lfsr_order(polynomial)
    g, t = start & (mask - 1), 1
    gen, tap, vec = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.int8)
    for i in range(n):
        gen[i], tap[i], vec[i] = g, t, (-1 if g & 1 else 1)
        g = lfsr_gen_step(g, polynomial, mask)
        t = lfsr_tap_step(t, polynomial, mask)
    return gen, tap, vec


def _lfsr_core(polynomial, start, x, flip_in, flip_out):
    """fastmat/LFSRCirculant.pyx:316-395: scatter rows to the generator-state addresses of a zeroed 2^order buffer
    (row 0 first, the remaining rows reversed when flip_in), Hadamard in the input's dtype, gather from the
    tap-state addresses (row 0 first, remaining rows reversed when flip_out)."""
    x2, was1d = _as2d(x)
    gen, tap, _ = lfsr_sequences(polynomial, start)
    n, order = gen.size, lfsr_order(polynomial)
    assert x2.shape[0] == n
    k = np.arange(n)
    flipped = np.where(k == 0, 0, n - k)
    data = np.zeros((1 << order, x2.shape[1]), dtype=x2.dtype)
    data[gen] = x2[flipped if flip_in else k]
    data = hadamard_forward(data, order)
    y = np.empty_like(x2)
    y[flipped if flip_out else k] = data[tap]
    return _ret(y, was1d)


def lfsr_circulant_forward(polynomial, start, x):
    """fastmat/LFSRCirculant.pyx:398-401."""
    return _lfsr_core(polynomial, start, x, True, False)


def lfsr_circulant_backward(polynomial, start, x):
    """fastmat/LFSRCirculant.pyx:403-406."""
    return _lfsr_core(polynomial, start, x, False, True)


def dense_lfsr_circulant(polynomial, start):
    """fastmat/LFSRCirculant.pyx:409-437: columns are the rolled output sequence."""
    _, _, vec = lfsr_sequences(polynomial, start)
    return np.stack([np.roll(vec, i) for i in range(vec.size)], axis=1)
