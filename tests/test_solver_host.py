"""CPU unit test of the statistical-equilibrium solver source that stat_equil_kernel runs on the GPU
(lightspinner_b200/csrc/mali_solve.h is host/device code; here it is compiled with g++ behind a tiny shim).

Checked against (a) the exact rational solution of the same fp64 system and (b) scipy.linalg.solve, i.e. the
reference's own call (rh_method.py:739), on the Gamma matrices of the reference fixture."""
import ctypes as C
import os
import subprocess
from fractions import Fraction

import numpy as np
import pytest

from helpers import load_golden

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('shim') / 'libsolver_shim.so')
    cxx = '/usr/bin/g++' if os.path.isfile('/usr/bin/g++') else 'g++'
    subprocess.check_call([cxx, '-O2', '-std=c++17', '-ffp-contract=off', '-shared', '-fPIC', '-o', out,
                           os.path.join(HERE, 'solver_shim.cpp')])
    L = C.CDLL(out)
    dp = C.POINTER(C.c_double)
    for f in (L.shim_solve8, L.shim_solve16, L.shim_solve_any):
        f.argtypes = [dp, C.c_int, C.c_int, C.c_double, dp]
    return L


def exact_solve(A, b):
    n = len(b)
    M = [[Fraction(float(A[i, j])) for j in range(n)] + [Fraction(float(b[i]))] for i in range(n)]
    for j in range(n):
        p = max(range(j, n), key=lambda i: abs(M[i][j]))
        M[j], M[p] = M[p], M[j]
        for i in range(j + 1, n):
            f = M[i][j] / M[j][j]
            for q in range(j, n + 1):
                M[i][q] -= f * M[j][q]
    x = [Fraction(0)] * n
    for i in range(n - 1, -1, -1):
        x[i] = (M[i][n] - sum(M[i][q] * x[q] for q in range(i + 1, n))) / M[i][i]
    return np.array([float(v) for v in x])


def solve_with(fn, G, iEl, nTot):
    NL = G.shape[0]
    Gc = np.ascontiguousarray(G, dtype=np.float64)
    x = np.zeros(NL)
    dp = C.POINTER(C.c_double)
    rc = fn(Gc.ctypes.data_as(dp), NL, iEl, float(nTot), x.ctypes.data_as(dp))
    return rc, x


def test_reference_systems_solved_to_the_rounded_exact_solution(shim, oracle):
    from scipy.linalg import solve
    p, r = load_golden('c1_falc_ca')
    oc = oracle.OracleContext(p)
    worst_exact, worst_scipy = 0.0, 0.0
    for it in range(1, 7):
        oc.formal_sol_gamma_matrices()
        if it > 3:
            G, n = oc.atom_Gamma(0), oc.atom_n(0)
            for k in range(0, oc.N, 3):
                iEl = int(np.argmax(n[:, k]))
                A = np.array(G[:, :, k])
                A[iEl, :] = 1.0
                b = np.zeros(6)
                b[iEl] = oc.nTotal[0, k]
                xe = exact_solve(A, b)
                for fn in (shim.shim_solve8, shim.shim_solve16):
                    rc, x = solve_with(fn, G[:, :, k], iEl, b[iEl])
                    assert rc == 0
                    worst_exact = max(worst_exact, np.max(np.abs(x - xe) / np.abs(xe)))
                    worst_scipy = max(worst_scipy, np.max(np.abs(x - solve(A, b)) / np.abs(xe)))
            oc.stat_equil()
    assert worst_exact <= 2.3e-16, worst_exact      # correctly rounded (<= 1 ulp)
    assert worst_scipy < 2e-10, worst_scipy          # LAPACK itself is ~8e-11 from exact here


def test_random_pivoting_systems(shim):
    rng = np.random.default_rng(5)
    for NL in (2, 3, 6, 8, 11, 16):
        fn = shim.shim_solve8 if NL <= 8 else shim.shim_solve16
        for trial in range(20):
            G = rng.normal(size=(NL, NL)) * np.exp(rng.normal(0, 4, size=(NL, NL)))
            iEl = int(rng.integers(0, NL))
            nTot = float(np.exp(rng.normal(30, 5)))
            A = G.copy()
            A[iEl, :] = 1.0
            b = np.zeros(NL)
            b[iEl] = nTot
            rc, x = solve_with(fn, G, iEl, nTot)
            assert rc == 0
            xe = exact_solve(A, b)
            assert np.max(np.abs(x - xe) / np.abs(xe)) < 1e-13, (NL, trial)


def test_singular_system_is_reported(shim):
    G = np.zeros((4, 4))
    G[1, :] = [1.0, 2.0, 3.0, 4.0]
    G[2, :] = [2.0, 4.0, 6.0, 8.0]
    rc, x = solve_with(shim.shim_solve8, G, 0, 1.0)
    assert rc == 1


def test_register_resident_form_returns_the_same_bits(shim):
    """solve_stat_equil_fixed<NL> (what the kernel runs for 2..8 levels) against the general form: identical bits,
    with and without row exchanges, and the same singular-system verdict."""
    rng = np.random.default_rng(11)
    for NL in (2, 3, 4, 5, 6, 7, 8):
        for trial in range(40):
            G = rng.normal(size=(NL, NL)) * np.exp(rng.normal(0, 4, size=(NL, NL)))
            if trial % 4 == 0:      # diagonally dominant: no exchanges
                G += np.diag(np.full(NL, 1e6))
            iEl = int(rng.integers(0, NL))
            nTot = float(np.exp(rng.normal(30, 5)))
            rc0, x0 = solve_with(shim.shim_solve8, G, iEl, nTot)
            rc1, x1 = solve_with(shim.shim_solve_any, G, iEl, nTot)
            assert rc0 == rc1 == 0
            assert np.array_equal(x0, x1), (NL, trial)
    G = np.zeros((4, 4))
    G[1, :] = [1.0, 2.0, 3.0, 4.0]
    G[2, :] = [2.0, 4.0, 6.0, 8.0]
    assert solve_with(shim.shim_solve_any, G, 0, 1.0)[0] == 1
