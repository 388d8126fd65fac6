"""CPU tests of the LFSRCirculant row: the numpy oracle and the native host logic
(fmb_lfsr_order / fmb_lfsr_period / fmb_lfsr_sequences) against fixtures frozen from the REAL reference
(oracle/make_golden_lfsr.py), and the scatter -> FWHT -> gather composition the class issues, run through the
host-emulated kernels (tests/emul, same CUDA source) with the class layer's own address vectors.  All integer: exact."""
import ctypes
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import fastmat_oracle as orc
from test_emulation_cpu import lib, apply, chk          # noqa: F401  (module-scoped fixture + helpers)

G = np.load(os.path.join(ROOT, 'tests', 'golden', 'golden_lfsr.npz'))
REGISTERS = [tuple(int(v) for v in r) for r in G['registers']]
DTYPES = ('int8', 'int32', 'int64', 'float64')


def tag(reg):
    return '%x_%x' % reg


@pytest.mark.parametrize('reg', REGISTERS, ids=tag)
def test_oracle_sequences_and_apply_match_reference(reg):
    t = tag(reg)
    gen, _, vec = orc.lfsr_sequences(*reg)
    assert orc.lfsr_period(*reg) == int(G[t + '_period'])
    assert np.array_equal(gen, G[t + '_states']) and np.array_equal(vec, G[t + '_vecC'])
    for dt in DTYPES:
        x = G['%s_%s_x' % (t, dt)]
        for fn, key in ((orc.lfsr_circulant_forward, 'fwd'), (orc.lfsr_circulant_backward, 'bwd')):
            y = fn(reg[0], reg[1], x)
            ref = G['%s_%s_%s' % (t, dt, key)]
            assert y.dtype == ref.dtype and np.array_equal(y, ref), (dt, key)
    if gen.size <= 127:        # the dense construction (fastmat/LFSRCirculant.pyx:409-437) agrees as well
        x = G[t + '_int64_x']
        assert np.array_equal(orc.dense_lfsr_circulant(*reg).astype(np.int64).dot(x), G[t + '_int64_fwd'])


def test_oracle_constructor_errors():
    for args, msg in (((0x19, 0), 'non-zero'), ((1, 1), 'order 1 to 31'), ((0x19, 0x10), 'non-zero')):
        with pytest.raises(ValueError, match=msg):
            orc.lfsr_period(*args)


@pytest.mark.parametrize('reg', REGISTERS, ids=tag)
def test_native_host_logic_matches_reference(reg):
    from fastmat_b200._lib import lib as clib
    from fastmat_b200.LFSRCirculant import lfsr_sequences
    t = tag(reg)
    n = int(clib.fmb_lfsr_period(*reg))
    assert n == int(G[t + '_period'])
    assert clib.fmb_lfsr_order(reg[0]) == orc.lfsr_order(reg[0])
    gen, tap, vec = lfsr_sequences(reg[0], reg[1], n)
    ogen, otap, ovec = orc.lfsr_sequences(*reg)
    assert np.array_equal(gen, G[t + '_states']) and np.array_equal(vec, G[t + '_vecC'])
    assert np.array_equal(tap, otap) and np.array_equal(gen, ogen) and np.array_equal(vec, ovec)


def test_native_host_logic_errors():
    from fastmat_b200._lib import lib as clib, FMB_ERR_VALUE
    for args, msg in (((0x19, 0), b'non-zero'), ((1, 1), b'order 1 to 31'), ((0x19, 0x10), b'non-zero')):
        assert clib.fmb_lfsr_period(*args) == FMB_ERR_VALUE and msg in clib.fmb_last_error()
    assert clib.fmb_lfsr_sequences(1, 1, 4, None, None, None) == FMB_ERR_VALUE
    assert clib.fmb_lfsr_sequences(0x19, 1, 0, None, None, None) == 0


@pytest.mark.parametrize('reg', REGISTERS, ids=tag)
def test_composition_emulated_bit_exact(lib, reg):          # noqa: F811
    """What LFSRCirculant._core launches: zero + scatter, Hadamard, gather, with the class layer's own address vectors
    (fastmat_b200.LFSRCirculant.lfsr_addresses).  The FWHT runs on the emulated kernel; the index kernels are not
    emulated (they are covered on the GPU by the Partial / Permutation parity tests), so numpy indexing with the
    semantics of fmb_partial_plan_create stands in for them here."""
    from fastmat_b200.LFSRCirculant import lfsr_addresses, lfsr_sequences
    t = tag(reg)
    n = int(G[t + '_period'])
    order = orc.lfsr_order(reg[0])
    gen, tap, _ = lfsr_sequences(reg[0], reg[1], n)
    scatter_flip, scatter, gather, gather_flip = lfsr_addresses(gen, tap)
    had = ctypes.c_void_p()
    chk(lib, lib.fmb_hadamard_plan_create(ctypes.byref(had), order))
    for dt in DTYPES:
        for lay in 'FC':
            x = np.asarray(G['%s_%s_x' % (t, dt)], order=lay)
            for sc, ga, key in ((scatter_flip, gather, 'fwd'), (scatter, gather_flip, 'bwd')):
                data = np.zeros((1 << order, x.shape[1]), dtype=x.dtype, order=lay)
                data[sc] = x                                    # partial plan, backward: y = 0; y[idx[r]] = x[r]
                data = apply(lib, had, 0, data, 1 << order, x.dtype)
                y = data[ga]                                    # partial plan, forward: y[r] = x[idx[r]]
                assert np.array_equal(y, G['%s_%s_%s' % (t, dt, key)]), (dt, lay, key)
    lib.fmb_plan_destroy(had)
