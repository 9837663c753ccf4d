"""CPU restatement of the per-column set-up that feeds the hot path.  TEST INFRASTRUCTURE ONLY:
only tests/ may import this module; the product (lightspinner_b200/) never does.

    lte_pops            atomic_set.py:105-145   (Debye-shifted Saha-Boltzmann populations)
    compute_collisions  rh_method.py:474-487 over collisional_rates.py:36-96 (Omega, CI, CE with scipy interp1d tables)
    v_broad             atomic_model.py:241-245

Plain numpy / scipy with the reference's own expressions and evaluation order, on plain arrays instead of the
reference's model objects.  Pinned bit for bit to the nStar, C and vBroad the unmodified reference produced
(tests/golden/*.npz; inputs in tests/golden/setup_inputs.npz) by tests/test_setup_oracle.py.
"""
import numpy as np
from scipy.interpolate import interp1d

# constants.py, digit for digit
HPlanck = 6.6260755E-34
KBoltzmann = 1.380658E-23
Amu = 1.6605402E-27
MElectron = 9.1093897E-31
QElectron = 1.60217733E-19
Epsilon0 = 8.854187817E-12
RBohr = 5.29177349E-11
ERydberg = 2.1798741E-18

OMEGA, CI, CE = 0, 1, 2


def lte_pops(E_SI, g, stage, temperature, ne, nTotal, debye=True):
    """atomic_set.py:105-145"""
    Nlevel = len(E_SI)
    c1 = (HPlanck / (2.0 * np.pi * MElectron)) * (HPlanck / KBoltzmann)
    c2 = 0.0
    nDebye = np.zeros(Nlevel)
    if debye:
        c2 = np.sqrt(8.0 * np.pi / KBoltzmann) * (QElectron**2 / (4.0 * np.pi * Epsilon0))**1.5
        for i in range(1, Nlevel):
            Z = int(stage[i])
            for m in range(1, int(stage[i]) - int(stage[0]) + 1):
                nDebye[i] += Z
                Z += 1
    dEion = c2 * np.sqrt(ne / temperature)
    cNe_T = 0.5 * ne * (c1 / temperature)**1.5
    total = np.ones(temperature.shape[0])
    nStar = np.zeros((Nlevel, temperature.shape[0]))
    for i in range(1, Nlevel):
        dE = E_SI[i] - E_SI[0]
        gi0 = g[i] / g[0]
        dZ = int(stage[i]) - int(stage[0])
        if debye:
            dE_kT = (dE - nDebye[i] * dEion) / (KBoltzmann * temperature)
        else:
            dE_kT = dE / (KBoltzmann * temperature)
        nst = gi0 * np.exp(-dE_kT)
        nStar[i, :] = nst
        nStar[i, :] /= cNe_T**dZ
        total += nStar[i]
    nStar[0] = nTotal / total
    for i in range(1, Nlevel):
        nStar[i] *= nStar[0]
    return nStar


def _interpolator(T, rates):
    """collisional_rates.py:15-19"""
    if len(rates) < 3:
        return interp1d(T, rates, fill_value=(rates[0], rates[-1]), bounds_error=False)
    return interp1d(T, rates, kind=3, fill_value=(rates[0], rates[-1]), bounds_error=False)


def compute_collisions(E_SI, g, coll, coll_T, coll_rates, temperature, ne, nStar):
    """rh_method.py:474-487: C[i, j] = rate from j to i, negatives clamped to zero.
    coll: [Ncoll, 4] = kind, i, j (i < j), number of table points; tables concatenated in coll_T / coll_rates."""
    Nlevel = len(E_SI)
    Cmat = np.zeros((Nlevel, Nlevel, temperature.shape[0]))
    o = 0
    for kind, i, j, n in coll:
        T, R = coll_T[o:o + n], coll_rates[o:o + n]
        o += n
        C = _interpolator(T, R)(temperature)
        if kind == OMEGA:       # collisional_rates.py:36-46
            C0 = ERydberg / np.sqrt(MElectron) * np.pi * RBohr**2 * np.sqrt(8.0 / (np.pi * KBoltzmann))
            Cdown = C0 * ne * C / (g[j] * np.sqrt(temperature))
            Cmat[i, j, :] += Cdown
            Cmat[j, i, :] += Cdown * nStar[j] / nStar[i]
        elif kind == CI:        # :64-72
            dE = E_SI[j] - E_SI[i]
            Cup = C * ne * np.exp(-dE / (KBoltzmann * temperature)) * np.sqrt(temperature)
            Cmat[j, i, :] += Cup
            Cmat[i, j, :] += Cup * nStar[i] / nStar[j]
        else:                   # CE :89-96
            gij = g[i] / g[j]
            Cdown = C * ne * gij * np.sqrt(temperature)
            Cmat[i, j, :] += Cdown
            Cmat[j, i, :] += Cdown * nStar[j] / nStar[i]
    Cmat[Cmat < 0.0] = 0.0
    return Cmat


def v_broad(weight, temperature, vturb):
    """atomic_model.py:241-245"""
    vTherm = 2.0 * KBoltzmann / (Amu * weight)
    return np.sqrt(vTherm * temperature + vturb**2)
