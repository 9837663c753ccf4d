"""State of the reference's EOS object for the device-side background (mali_model_set_eos / mali_background):
partition-function tables, abundances and the handful of derived scalars a witt.witt() instance holds
(witt.py:152-195), plus the constants Background.compute_background_eos uses around it (background.py:11, 33).

    et = EosTables.from_witt(witt.witt(), atomicTable.weightPerH)      # reference objects
    et = EosTables.from_arrays(npz)                                    # tests/golden/eos.npz
"""
import ctypes as C

import numpy as np

from . import _capi

# constants.py / witt.py:41-49, digit for digit
Amu = 1.6605402E-27
CM_TO_M = 1.0E-02
QElectron = 1.60217733E-19
Epsilon0 = 8.854187817E-12
MElectron = 9.1093897E-31
CLight = 2.99792458E+08
W_PI, W_ME, W_BK, W_HH = 3.14159265358979323846, 9.10938188E-28, 1.3806488E-16, 6.62606957E-27
NCONTR = 28


class EosTables:
    def __init__(self, tpf, pf, eion, stage_off, abund, avw, rho_from_H, ab_others, weightPerH, prec=1.e-5):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        self.tpf, self.pf, self.eion, self.abund = f(tpf), f(pf), f(eion), f(abund)
        self.stage_off = np.ascontiguousarray(stage_off, dtype=np.int32)
        if self.stage_off.shape[0] != NCONTR + 1 or self.abund.shape[0] != 99:
            raise ValueError('EOS tables: 28 elements with all their stages and 99 abundances are expected')
        self.avw, self.rho_from_H, self.ab_others = float(avw), float(rho_from_H), float(ab_others)
        self.prec = float(prec)
        self.saha_fac = ((2.0 * W_PI * W_ME * W_BK) / (W_HH * W_HH))**1.5                      # witt.py:52
        self.amu_wph = Amu * float(weightPerH)                                                  # background.py:33
        self.cm3 = CM_TO_M**3
        self.thomson = 8.0 * np.pi / 3.0 * (QElectron / (np.sqrt(4.0 * np.pi * Epsilon0) *
                                                         (np.sqrt(MElectron) * CLight)))**4      # background.py:11

    @classmethod
    def from_arrays(cls, z):
        return cls(z['tpf'], z['pf'], z['eion'], z['stage_off'], z['ABUND'], z['avw'], z['rho_from_H'], z['ab_others'],
                   z['weightPerH'])

    @classmethod
    def from_witt(cls, eos, weightPerH):
        """From a live witt.witt() instance of the reference."""
        pf = np.concatenate([np.asarray(eos.el[i].pf) for i in range(NCONTR)], axis=0)
        eion = np.concatenate([np.asarray(eos.el[i].eion) for i in range(NCONTR)])
        off = np.concatenate([[0], np.cumsum([eos.el[i].nstage for i in range(NCONTR)])])
        return cls(eos.tpf, pf, eion, off, eos.ABUND, eos.avw, eos.rho_from_H, eos.ab_others, weightPerH, eos.prec)

    def desc(self):
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        return _capi.EosDesc(int(self.tpf.shape[0]), dp(self.tpf), dp(self.pf), dp(self.eion),
                             self.stage_off.ctypes.data_as(C.POINTER(C.c_int32)), dp(self.abund), self.avw,
                             self.rho_from_H, self.ab_others, self.saha_fac, self.prec, self.amu_wph, self.cm3,
                             float(self.thomson))
