#!/usr/bin/env python
"""Generate tests/golden/golden_lfsr.npz from the REAL reference's LFSRCirculant (TEST INFRASTRUCTURE ONLY).

    bash oracle/build_ref.sh && python oracle/make_golden_lfsr.py

Registers: the two of the reference's own test (fastmat/LFSRCirculant.pyx:441-447: 0x19 / 0x17 with starts 0xD / 0x1),
a maximum-length order-10 register, and a non-maximum-length order-12 one (zero-filled Hadamard rows, :327-330).
All integer: the outputs are exact.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '_ref'))
import fastmat as fm                                    # noqa: E402  (the real reference)

OUT = os.path.join(HERE, '..', 'tests', 'golden', 'golden_lfsr.npz')
rng = np.random.default_rng(31337)
REGISTERS = [(0x19, 0xD), (0x19, 0x1), (0x17, 0xD), (0x17, 0x1), (0x409, 0x2A5), (0x1003, 0x7)]
arrs = {'registers': np.array(REGISTERS, dtype=np.int64)}
for poly, start in REGISTERS:
    L = fm.LFSRCirculant(poly, start)
    tag = '%x_%x' % (poly, start)
    n = L.numRows
    arrs[tag + '_period'] = np.int64(L.period)
    arrs[tag + '_states'] = np.array(L.states)
    arrs[tag + '_vecC'] = np.array(L.vecC)
    for dt in ('int8', 'int32', 'int64', 'float64'):
        x = rng.integers(-3, 3, size=(n, 5), endpoint=True).astype(dt)
        arrs['%s_%s_x' % (tag, dt)] = x
        arrs['%s_%s_fwd' % (tag, dt)] = L.forward(x)
        arrs['%s_%s_bwd' % (tag, dt)] = L.backward(x)
    if n <= 1023:
        ref = L.reference()
        x = arrs[tag + '_int64_x']
        assert np.array_equal(ref.astype(np.int64).dot(x), arrs[tag + '_int64_fwd'])
        assert np.array_equal(ref.astype(np.int64).T.dot(x), arrs[tag + '_int64_bwd'])
np.savez_compressed(OUT, **arrs)
print('wrote', OUT, os.path.getsize(OUT), 'bytes')
