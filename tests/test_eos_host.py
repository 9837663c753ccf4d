"""lightspinner_b200/csrc/mali_eos.h (Wittmann EOS + background opacity, the source the device kernels compile) built
for the host and checked against the numbers the unmodified reference produced (tests/golden/eos.npz: one instrumented
witt() instance, tests/golden/make_golden.py eos): gas / electron pressure, the 17 background partial densities, the
continuum opacity on the CaII + H wavelength grid and at 5000 A, for FALC, a config-4 jitter column and a
response-function column."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)


class Tables(C.Structure):
    _fields_ = [('npf', C.c_int32), ('tpf', _dp), ('pf', _dp), ('eion', _dp), ('stageOff', C.POINTER(C.c_int32)),
                ('abund', _dp), ('avw', C.c_double), ('rho_from_H', C.c_double), ('ab_others', C.c_double),
                ('saha_fac', C.c_double), ('prec', C.c_double)]


def eos_tables(z):
    """(ctypes struct, keep-alive arrays) from the witt() state stored in eos.npz."""
    PI, ME, BK, HH = 3.14159265358979323846, 9.10938188E-28, 1.3806488E-16, 6.62606957E-27
    keep = [np.ascontiguousarray(z[k], dtype=np.float64) for k in ('tpf', 'pf', 'eion', 'ABUND')]
    so = np.ascontiguousarray(z['stage_off'], dtype=np.int32)
    keep.append(so)
    t = Tables(int(z['tpf'].shape[0]), keep[0].ctypes.data_as(_dp), keep[1].ctypes.data_as(_dp), keep[2].ctypes.data_as(_dp),
               so.ctypes.data_as(C.POINTER(C.c_int32)), keep[3].ctypes.data_as(_dp), float(z['avw']), float(z['rho_from_H']),
               float(z['ab_others']), ((2.0 * PI * ME * BK) / (HH * HH))**1.5, 1.e-5)
    return t, keep


@pytest.fixture(scope='module')
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('eshim') / 'libeos_shim.so')
    cxx = '/usr/bin/g++' if os.path.isfile('/usr/bin/g++') else 'g++'
    subprocess.check_call([cxx, '-O2', '-std=c++17', '-ffp-contract=off', '-shared', '-fPIC', '-o', out,
                           os.path.join(HERE, 'eos_shim.cpp')])
    L = C.CDLL(out)
    L.shim_eos.argtypes = [C.POINTER(Tables), C.c_int, _dp, _dp, _dp, _dp]
    L.shim_partials.argtypes = [C.POINTER(Tables), C.c_int, _dp, _dp, _dp, _dp]
    L.shim_cont_opacity.argtypes = [C.POINTER(Tables), C.c_double, C.c_double, C.c_double, C.c_int, _dp, _dp]
    return L


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


@pytest.mark.parametrize('case', ['falc', 'jitter0', 'rf_k40p'])
def test_eos_and_opacity_match_the_reference(shim, case):
    z = np.load(os.path.join(HERE, 'golden', 'eos.npz'))
    E, keep = eos_tables(z)
    T = np.ascontiguousarray(z[case + '_T'])
    rho = np.ascontiguousarray(z[case + '_rho'])
    N = T.shape[0]
    pgas, pe = np.zeros(N), np.zeros(N)
    P = lambda a: a.ctypes.data_as(_dp)
    shim.shim_eos(C.byref(E), N, P(T), P(rho), P(pgas), P(pe))
    e_pg, e_pe = rel(pgas, z[case + '_pgas']), rel(pe, z[case + '_pe'])
    # the remaining checks take the reference's pressures as inputs, so that each stage is judged on its own
    rp, re = np.ascontiguousarray(z[case + '_pgas']), np.ascontiguousarray(z[case + '_pe'])
    parts = np.zeros((N, 17))
    shim.shim_partials(C.byref(E), N, P(T), P(rp), P(re), P(parts))
    e_parts = rel(parts, z[case + '_partials'])
    wav = np.ascontiguousarray(z['wavelength'] * 10)
    chi = np.zeros((wav.shape[0], N))
    col = np.zeros(wav.shape[0])
    chi_c = np.zeros(N)
    w5 = np.array([5000.0])
    one = np.zeros(1)
    for k in range(N):
        shim.shim_cont_opacity(C.byref(E), T[k], rp[k], re[k], wav.shape[0], P(wav), P(col))
        chi[:, k] = col / 1.0E-02
        shim.shim_cont_opacity(C.byref(E), T[k], rp[k], re[k], 1, P(w5), P(one))
        chi_c[k] = one[0] / 1.0E-02
    e_chi, e_c = rel(chi, z[case + '_chi']), rel(chi_c, z[case + '_chi_c'])
    print('%s: pgas %.2e pe %.2e partials %.2e chi %.2e chi_c %.2e' % (case, e_pg, e_pe, e_parts, e_chi, e_c))
    assert e_pg < 1e-12 and e_pe < 1e-12
    assert e_parts < 1e-12
    assert e_chi < 1e-12 and e_c < 1e-12
