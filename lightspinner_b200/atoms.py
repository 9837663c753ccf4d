"""Model-level atom data for the device-side column set-up (mali_model_set_atoms / mali_setup_columns): what the
reference's lte_pops (atomic_set.py:105-145), collisional-rate terms (collisional_rates.py:36-96) and v_broad
(atomic_model.py:241-245) read from its atomic-model objects, flattened once per model.

    at = AtomTables.from_models([atom.atomicModel for atom in ctx.activeAtoms], atomicTable)   # reference objects
    at = AtomTables.from_arrays([dict(E_SI=..., g=..., stage=..., weight=..., coll=..., coll_T=..., coll_rates=...), ...])

Everything here is iteration- and column-invariant host work; the per-column arithmetic runs on the GPU.
"""
import ctypes as C

import numpy as np

from . import _capi

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
_KINDS = {'Omega': OMEGA, 'CI': CI, 'CE': CE}


def _interpolant(T, rates):
    """The interpolant of collisional_rates.py:15-19 (scipy interp1d: cubic not-a-knot B-spline, linear below 3
    points) as (knots, coefficients, cubic): the B-spline's knot vector and coefficients, which the device evaluates
    with the same de Boor recurrence scipy uses, or the table itself for the linear form."""
    T = np.ascontiguousarray(T, dtype=np.float64)
    rates = np.ascontiguousarray(rates, dtype=np.float64)
    if len(T) < 2:
        raise ValueError('a collisional-rate table needs at least 2 temperatures')
    if len(T) < 3:
        return T, rates, 0
    from scipy.interpolate import make_interp_spline
    spl = make_interp_spline(T, rates, k=3, check_finite=False)     # what interp1d(kind=3) builds
    return np.ascontiguousarray(spl.t, dtype=np.float64), np.ascontiguousarray(spl.c, dtype=np.float64), 1


class AtomTables:
    def __init__(self, atoms):
        """atoms: list (the model's atom order) of dicts with E_SI, g, stage [Nlevel], weight (atomic weight),
        coll [Ncoll, 4] = kind, i, j, table points; coll_T, coll_rates concatenated tables."""
        self.Natom = len(atoms)
        self.Nlevel = np.array([len(a['E_SI']) for a in atoms], dtype=np.int32)
        dE, gi0, dZ, nDebye, g, vTherm = [], [], [], [], [], []
        coll, knots, coef, fill, par = [], [], [], [], []
        nk = ncf = 0
        for ia, a in enumerate(atoms):
            E = np.asarray(a['E_SI'], dtype=np.float64)
            gg = np.asarray(a['g'], dtype=np.float64)
            st = np.asarray(a['stage'], dtype=np.int64)
            dE.append(E - E[0])                                   # atomic_set.py:131
            gi0.append(gg / gg[0])                                # :132
            dZ.append((st - st[0]).astype(np.int32))              # :133
            nd = np.zeros(len(E))
            for i in range(1, len(E)):                            # :113-119
                Z = int(st[i])
                for m in range(1, int(st[i]) - int(st[0]) + 1):
                    nd[i] += Z
                    Z += 1
            nDebye.append(nd)
            g.append(gg)
            vTherm.append(2.0 * KBoltzmann / (Amu * float(a['weight'])))    # atomic_model.py:242
            o = 0
            for kind, i, j, n in np.asarray(a['coll'], dtype=np.int64).reshape(-1, 4):
                i, j = min(i, j), max(i, j)                       # collisional_rates.py:29-31
                T = np.asarray(a['coll_T'], dtype=np.float64)[o:o + n]
                R = np.asarray(a['coll_rates'], dtype=np.float64)[o:o + n]
                o += n
                t, cf, cubic = _interpolant(T, R)
                coll.append([ia, kind, i, j, n, nk, ncf, cubic])
                knots.append(t)
                coef.append(cf)
                fill.append([R[0], R[-1]])
                nk += len(t)
                ncf += len(cf)
                if kind == OMEGA:                                 # :35
                    par.append(ERydberg / np.sqrt(MElectron) * np.pi * RBohr**2 * np.sqrt(8.0 / (np.pi * KBoltzmann)))
                elif kind == CI:                                  # :62
                    par.append(E[j] - E[i])
                else:                                             # :87
                    par.append(gg[i] / gg[j])
        cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if xs else np.zeros(0), dtype=dt)
        self.dE, self.gi0, self.nDebye, self.g = (cat(x, np.float64) for x in (dE, gi0, nDebye, g))
        self.dZ = cat(dZ, np.int32)
        self.vTherm = np.ascontiguousarray(vTherm, dtype=np.float64)
        self.coll = np.ascontiguousarray(np.array(coll, dtype=np.int32).reshape(-1, 8))
        self.knots = cat(knots, np.float64)
        self.coef = cat(coef, np.float64)
        self.fill = np.ascontiguousarray(np.array(fill, dtype=np.float64).reshape(-1, 2))
        self.par = np.ascontiguousarray(par, dtype=np.float64)
        self.c1 = (HPlanck / (2.0 * np.pi * MElectron)) * (HPlanck / KBoltzmann)                              # atomic_set.py:107
        self.c2 = np.sqrt(8.0 * np.pi / KBoltzmann) * (QElectron**2 / (4.0 * np.pi * Epsilon0))**1.5          # :111

    @classmethod
    def from_arrays(cls, atoms):
        return cls(atoms)

    @classmethod
    def from_models(cls, models, atomicTable):
        """From the reference's AtomicModel objects (the model's atom order) and its AtomicTable (atomic weights)."""
        atoms = []
        for m in models:
            meta, T, R = [], [], []
            for c in m.collisions:
                kind = _KINDS.get(type(c).__name__)
                if kind is None:
                    raise ValueError('collisional-rate term %s has no device form' % type(c).__name__)
                meta.append([kind, c.i, c.j, len(c.temperature)])
                T.append(np.asarray(c.temperature, dtype=np.float64))
                R.append(np.asarray(c.rates, dtype=np.float64))
            atoms.append(dict(E_SI=[l.E_SI for l in m.levels], g=[l.g for l in m.levels],
                              stage=[l.stage for l in m.levels], weight=atomicTable[m.name].weight,
                              coll=np.array(meta, dtype=np.int32).reshape(-1, 4),
                              coll_T=np.concatenate(T) if T else np.zeros(0),
                              coll_rates=np.concatenate(R) if R else np.zeros(0)))
        return cls(atoms)

    def desc(self):
        """ctypes mali_atom_desc viewing this object's arrays (keep `self` alive while it is used)."""
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        return _capi.AtomDesc(self.Natom, ip(self.Nlevel), dp(self.dE), dp(self.gi0), ip(self.dZ), dp(self.nDebye),
                              dp(self.g), dp(self.vTherm), float(self.c1), float(self.c2), int(self.coll.shape[0]),
                              ip(self.coll), dp(self.knots), dp(self.coef), dp(self.fill), dp(self.par))
