"""Shared test helpers: golden-fixture loading and comparison metrics."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    """Returns (problem dict, results dict) of tests/golden/<name>.npz."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    p = {k[2:]: z[k] for k in z.files if k.startswith('p_')}
    r = {k[2:]: z[k] for k in z.files if k.startswith('r_')}
    for k in ('Nspace', 'Nrays', 'Nspect'):
        p[k] = int(p[k])
    return p, r


def load_units():
    return np.load(os.path.join(GOLDEN, 'units.npz'))


def relerr(a, b):
    """max |a-b| / |b| over entries where b != 0 (plus absolute check where b == 0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    nz = b != 0
    e = 0.0
    if nz.any():
        e = float(np.max(np.abs(a[nz] - b[nz]) / np.abs(b[nz])))
    if (~nz).any():
        e = max(e, float(np.max(np.abs(a[~nz]))))
    return e
