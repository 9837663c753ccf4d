"""mali_voigt.h (the Voigt function of the device compute_phi) against scipy.special.wofz, the function the
reference calls (utils.py:13-15).  Host build of the shared host/device source; the GPU test compares whole
profiles."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
from scipy.special import wofz

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('vshim') / 'libvoigt_shim.so')
    cxx = '/usr/bin/g++' if os.path.isfile('/usr/bin/g++') else 'g++'
    subprocess.check_call([cxx, '-O2', '-std=c++17', '-shared', '-fPIC', '-o', out, os.path.join(HERE, 'voigt_shim.cpp')])
    L = C.CDLL(out)
    dp = C.POINTER(C.c_double)
    L.shim_voigt.argtypes = [C.c_int, dp, dp, dp]
    return L


def voigt(shim, a, v):
    a = np.ascontiguousarray(a, dtype=np.float64)
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.empty_like(v)
    dp = C.POINTER(C.c_double)
    shim.shim_voigt(v.size, a.ctypes.data_as(dp), v.ctypes.data_as(dp), out.ctypes.data_as(dp))
    return out


def test_voigt_matches_wofz_over_the_damping_and_frequency_range(shim):
    rng = np.random.default_rng(7)
    worst = 0.0
    for a in (1e-7, 1e-5, 1e-4, 1e-3, 1e-2, 0.1, 0.5, 1.0, 3.0, 6.0, 6.28, 6.3, 10.0, 30.0):
        v = np.concatenate([np.linspace(0, 12, 20001), np.logspace(0, 4, 4001), rng.uniform(0, 8, 20000),
                            np.arange(0, 60) * 0.25, -rng.uniform(0, 30, 1000)])
        ref = wofz(v + 1j * a).real
        got = voigt(shim, np.full_like(v, a), v)
        worst = max(worst, float(np.max(np.abs(got - ref) / ref)))
    assert worst < 2e-13, worst      # wofz itself is only good to ~1.2e-13 (next test)


def test_voigt_random_pairs(shim):
    rng = np.random.default_rng(8)
    a = 10.0**rng.uniform(-7, 1.4, 200000)
    v = rng.uniform(-1, 1, 200000) * 10.0**rng.uniform(-3, 3.5, 200000)
    ref = wofz(v + 1j * a).real
    got = voigt(shim, a, v)
    assert float(np.max(np.abs(got - ref) / ref)) < 2e-13


def test_voigt_against_40_digit_arithmetic(shim):
    """Ground truth from mpmath: w(z) = exp(-z^2) erfc(-i z).  mali_voigt.h is good to a few ulp; scipy's wofz (what
    the reference calls) to ~1e-13 -- which is therefore the floor of any profile comparison with the reference."""
    mp = pytest.importorskip('mpmath')
    mp.mp.dps = 40
    rng = np.random.default_rng(3)
    # random pairs, the region where wofz switches algorithm, and both sides of mali_voigt.h's |z| = 16 switch
    a = np.concatenate([10.0**rng.uniform(-7, 1.4, 600), rng.uniform(0.09, 0.17, 150), 10.0**rng.uniform(-7, 1.2, 200),
                        rng.uniform(15, 17, 100)])
    v = np.concatenate([rng.uniform(0, 1, 600) * 10.0**rng.uniform(-3, 3.5, 600), rng.uniform(5.9, 6.2, 150),
                        rng.uniform(15.9, 18, 200), rng.uniform(0, 6, 100)])
    got = voigt(shim, a, v)

    def exact(x, y):
        z = mp.mpc(x, y)
        return float(mp.re(mp.exp(-z * z) * mp.erfc(-1j * z)))

    ex = np.array([exact(float(x), float(y)) for x, y in zip(v, a)])
    assert float(np.max(np.abs(got - ex) / ex)) < 4e-15
    assert float(np.max(np.abs(wofz(v + 1j * a).real - ex) / ex)) < 5e-13
