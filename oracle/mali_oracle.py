"""ctypes front-end of the CPU oracle (oracle/mali_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (lightspinner_b200/) never does.

A *problem* is a plain dict of numpy arrays in the reference's own layouts:

model part (shared by columns)
    Nspace, Nrays, Nspect      ints
    wavelength [Nspect], muz [Nrays], wmu [Nrays]
    Nlevel [Natom]             levels per active atom (activeAtoms order, rh_method.py:558-560)
    trans  [Ntrans, 6]         atom, i, j, isLine, Nblue, Nlambda -- atom.trans order (rh_method.py:398-405)
    linepar [Ntrans, 4]        Aji, Bji, Bij, lambda0 (zeros for continua)
    alpha  [sum Nlambda]       continuum cross-sections, concatenated with offsets toff (zeros for lines)
column part
    height, temperature [Nspace]
    bg_chi, bg_eta, bg_sca [Nspect, Nspace]                     background.py:37-51
    nStar [sumNlevel, Nspace], nTotal [Natom, Nspace]
    C [sum Nlevel^2, Nspace]   per atom C[i, j, k] flattened (rh_method.py:474-487)
    n [sumNlevel, Nspace]      current populations
    phi  1-D concat of each line's phi[Nlambda, Nrays, 2, Nspace] (rh_method.py:224), phioff [Ntrans]
    wphi [Ntrans, Nspace]      (rows of continua are zero)
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_build', 'libmali_oracle.so')
_lib = None

# constants.py:1-4,17
CLight = 2.99792458E+08
HPlanck = 6.6260755E-34
HC = HPlanck * CLight
KBoltzmann = 1.380658E-23
NM_TO_M = 1.0E-09

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)


class _Model(C.Structure):
    _fields_ = [('Nspace', C.c_int), ('Nrays', C.c_int), ('Nspect', C.c_int), ('Natom', C.c_int),
                ('Ntrans', C.c_int), ('Nlevel', _ip), ('trans', _ip), ('toff', _ip), ('wavelength', _dp),
                ('muz', _dp), ('wmu', _dp), ('lineconst', _dp), ('wlambda', _dp), ('alpha', _dp),
                ('twohc_l3', _dp), ('wlacont', _dp)]


class _Column(C.Structure):
    _fields_ = [('height', _dp), ('temperature', _dp), ('bg_chi', _dp), ('bg_eta', _dp), ('bg_sca', _dp),
                ('nTotal', _dp), ('C', _dp), ('phi', _dp), ('phioff', _lp), ('wphi', _dp), ('gijcont', _dp),
                ('n', _dp), ('J', _dp), ('I', _dp), ('Gamma', _dp)]


def build(force=False):
    """Compile the C restatement (gcc).  Building the checker is not using it."""
    src = os.path.join(_HERE, 'mali_oracle.c')
    if (not force and os.path.isfile(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src),
                                                   os.path.getmtime(os.path.join(_HERE, 'mali_oracle.h')))):
        return _LIB_PATH
    subprocess.check_call(['make', '-C', _HERE, '-s'] + (['-B'] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.lso_w2.argtypes = [C.c_double, _dp]
        L.lso_planck.restype = C.c_double
        L.lso_planck.argtypes = [C.c_double, C.c_double]
        L.lso_piecewise_1d_impl.argtypes = [C.c_double, C.c_int, C.c_double, _dp, _dp, _dp, C.c_int, _dp, _dp]
        L.lso_piecewise_linear_1d.argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_int, C.c_double, _dp, _dp, _dp, _dp]
        L.lso_uv.argtypes = [C.POINTER(_Model), C.POINTER(_Column), C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp,
                             _dp, _dp]
        L.lso_formal_sol_gamma_matrices.restype = C.c_double
        L.lso_formal_sol_gamma_matrices.argtypes = [C.POINTER(_Model), C.POINTER(_Column)]
        L.lso_stat_equil.restype = C.c_double
        L.lso_stat_equil.argtypes = [C.POINTER(_Model), C.POINTER(_Column), _ip]
        L.lso_iterate_batch.argtypes = [C.POINTER(_Model), C.POINTER(_Column), C.c_int, C.c_int, C.c_int, _dp, _dp]
        L.lso_max_threads.restype = C.c_int
        L.lso_set_threads.argtypes = [C.c_int]
        L.lso_set_threads.restype = None
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ci(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def w2(dtau):
    out = np.zeros(2)
    lib().lso_w2(float(dtau), _d(out))
    return out


def planck(temp, wav):
    return lib().lso_planck(float(temp), float(wav))


def piecewise_linear_1d(height, temperature, muz, toFrom, wav, chi, S):
    """formal_solver.py:144-212 -> (I, PsiStar)"""
    height, temperature, chi, S = map(_c64, (height, temperature, chi, S))
    N = chi.shape[0]
    I = np.zeros(N)
    Psi = np.zeros(N)
    lib().lso_piecewise_linear_1d(_d(height), _d(temperature), N, float(muz), int(bool(toFrom)), float(wav),
                                  _d(chi), _d(S), _d(I), _d(Psi))
    return I, Psi


def toff_of(trans):
    return np.concatenate([[0], np.cumsum(trans[:, 5])]).astype(np.int32)


def model_tables(p):
    """Iteration-invariant per-model tables, evaluated with the reference's own scalar numpy expressions.

    lineconst : rh_method.py:268,279,281 (hc_4pi*Bij, Aji/Bji) and :450 (Bji/Bij)
    wlambda   : rh_method.py:178-189
    twohc_l3  : rh_method.py:286
    wlacont   : rh_method.py:455
    """
    trans = np.asarray(p['trans'])
    wavelength = np.asarray(p['wavelength'], dtype=np.float64)
    toff = toff_of(trans)
    ntot = int(toff[-1])
    lineconst = np.zeros((trans.shape[0], 3))
    wlambda = np.zeros(ntot)
    twohc = np.zeros(ntot)
    wlacont = np.zeros(ntot)
    hc_4pi = 0.25 * HC / np.pi
    w_given = p.get('wlambda_table')   # weights formed on a larger grid (wavelength-sharded column)
    for t, (atom, i, j, isLine, Nblue, Nlam) in enumerate(trans):
        wl = wavelength[Nblue:Nblue + Nlam]
        if isLine:
            Aji, Bji, Bij, lambda0 = (np.float64(x) for x in p['linepar'][t])
            lineconst[t] = (hc_4pi * Bij, Aji / Bji, Bji / Bij)
            dopplerWidth = CLight / lambda0
        else:
            dopplerWidth = 1.0
        for lt in range(Nlam):
            if w_given is not None:
                w = np.float64(w_given[toff[t] + lt])
            elif lt == 0:
                w = 0.5 * (wl[1] - wl[0]) * dopplerWidth
            elif lt == Nlam - 1:
                w = 0.5 * (wl[-1] - wl[-2]) * dopplerWidth
            else:
                w = 0.5 * (wl[lt + 1] - wl[lt - 1]) * dopplerWidth
            wlambda[toff[t] + lt] = w
            if not isLine:
                twohc[toff[t] + lt] = 2.0 * HC / (NM_TO_M * wl[lt])**3
                wlacont[toff[t] + lt] = w / wl[lt] / HPlanck
    return dict(toff=toff, lineconst=lineconst, wlambda=wlambda, twohc_l3=twohc, wlacont=wlacont)


def column_tables(p):
    """Continuum g_ij[(t, lt), k] exactly as rh_method.py:453-454 (numpy vector exp over depth)."""
    trans = np.asarray(p['trans'])
    toff = toff_of(trans)
    N = int(p['Nspace'])
    lvloff = np.concatenate([[0], np.cumsum(p['Nlevel'])])
    gij = np.zeros((int(toff[-1]), N))
    hc_k = HC / (KBoltzmann * NM_TO_M)
    temperature = _c64(p['temperature'])
    nStar = _c64(p['nStar'])
    wavelength = np.asarray(p['wavelength'], dtype=np.float64)
    for t, (atom, i, j, isLine, Nblue, Nlam) in enumerate(trans):
        if isLine:
            continue
        ni = np.ascontiguousarray(nStar[lvloff[atom] + i])
        nj = np.ascontiguousarray(nStar[lvloff[atom] + j])
        for lt in range(Nlam):
            gij[toff[t] + lt, :] = ni / nj * np.exp(-hc_k / wavelength[Nblue + lt] / temperature)
    return dict(gijcont=gij)


class OracleContext:
    """Mirror of rh_method.Context for one column, driven by the C restatement.

    stat_equil(use_scipy=True) performs the very call the reference makes (scipy.linalg.solve,
    rh_method.py:739) so that a lock-step run can be bit-identical to the reference;
    use_scipy=False uses the oracle's own LU (what the timed CPU baseline runs).
    """

    def __init__(self, problem):
        p = problem
        self.p = p
        self.N = int(p['Nspace'])
        self.Nrays = int(p['Nrays'])
        self.Nspect = int(p['Nspect'])
        self.Nlevel = _ci(p['Nlevel'])
        self.Natom = self.Nlevel.shape[0]
        self.trans = _ci(p['trans'])
        self.Ntrans = self.trans.shape[0]
        self.lvloff = np.concatenate([[0], np.cumsum(self.Nlevel)]).astype(int)
        self.g2off = np.concatenate([[0], np.cumsum(self.Nlevel.astype(int)**2)]).astype(int)
        mt = model_tables(p)
        ct = column_tables(p)
        self._keep = dict(
            wavelength=_c64(p['wavelength']), muz=_c64(p['muz']), wmu=_c64(p['wmu']),
            toff=_ci(mt['toff']), lineconst=_c64(mt['lineconst']), wlambda=_c64(mt['wlambda']),
            alpha=_c64(p['alpha']), twohc=_c64(mt['twohc_l3']), wlacont=_c64(mt['wlacont']),
            height=_c64(p['height']), temperature=_c64(p['temperature']), bg_chi=_c64(p['bg_chi']),
            bg_eta=_c64(p['bg_eta']), bg_sca=_c64(p['bg_sca']), nTotal=_c64(p['nTotal']), C=_c64(p['C']),
            phi=_c64(p['phi']), phioff=np.ascontiguousarray(p['phioff'], dtype=np.int64),
            wphi=_c64(p['wphi']), gijcont=_c64(ct['gijcont']))
        k = self._keep
        self.n = np.array(p['n'], dtype=np.float64, order='C', copy=True)
        self.J = np.zeros((self.Nspect, self.N))
        self.I = np.zeros((self.Nspect, self.Nrays))
        self.Gamma = np.zeros((int(self.g2off[-1]), self.N))
        self.nTotal = k['nTotal']
        self._m = _Model(self.N, self.Nrays, self.Nspect, self.Natom, self.Ntrans,
                         self.Nlevel.ctypes.data_as(_ip), self.trans.ctypes.data_as(_ip),
                         k['toff'].ctypes.data_as(_ip), _d(k['wavelength']), _d(k['muz']), _d(k['wmu']),
                         _d(k['lineconst']), _d(k['wlambda']), _d(k['alpha']), _d(k['twohc']), _d(k['wlacont']))
        self._c = _Column(_d(k['height']), _d(k['temperature']), _d(k['bg_chi']), _d(k['bg_eta']),
                          _d(k['bg_sca']), _d(k['nTotal']), _d(k['C']), _d(k['phi']),
                          k['phioff'].ctypes.data_as(_lp), _d(k['wphi']), _d(k['gijcont']), _d(self.n),
                          _d(self.J), _d(self.I), _d(self.Gamma))

    # -- accessors in the reference's shapes
    def atom_n(self, a):
        return self.n[self.lvloff[a]:self.lvloff[a + 1]]

    def atom_Gamma(self, a):
        NL = int(self.Nlevel[a])
        return self.Gamma[self.g2off[a]:self.g2off[a + 1]].reshape(NL, NL, self.N)

    def formal_sol_gamma_matrices(self):
        return lib().lso_formal_sol_gamma_matrices(C.byref(self._m), C.byref(self._c))

    def uv(self, t, la, mu, toFrom, gij):
        gij = _c64(gij)
        U, Vij, Vji = np.zeros(self.N), np.zeros(self.N), np.zeros(self.N)
        lib().lso_uv(C.byref(self._m), C.byref(self._c), t, la, mu, int(toFrom), _d(gij), _d(U), _d(Vij), _d(Vji))
        return U, Vij, Vji

    def stat_equil(self, use_scipy=True):
        if not use_scipy:
            sing = C.c_int(0)
            r = lib().lso_stat_equil(C.byref(self._m), C.byref(self._c), C.byref(sing))
            if sing.value:
                raise np.linalg.LinAlgError('singular statistical-equilibrium system')
            return r
        from scipy.linalg import solve
        maxRelChange = 0.0
        for a in range(self.Natom):
            n = self.atom_n(a)
            G = self.atom_Gamma(a)
            NL = int(self.Nlevel[a])
            for k in range(self.N):
                iEliminate = np.argmax(n[:, k])
                Gamma = np.copy(G[:, :, k])
                Gamma[iEliminate, :] = 1.0
                nk = np.zeros(NL)
                nk[iEliminate] = self.nTotal[a, k]
                nOld = np.copy(n[:, k])
                nNew = solve(Gamma, nk)
                change = np.abs(1.0 - nOld / nNew)
                maxRelChange = max(maxRelChange, change.max())
                n[:, k] = nNew
        return maxRelChange


def iterate_batch(problems, niter, start_iter=0):
    """lso_iterate_batch over a list of OracleContext objects sharing one model.  Returns (dJ, dPops)."""
    ctxs = list(problems)
    ncol = len(ctxs)
    cols = (_Column * ncol)(*[c._c for c in ctxs])
    dJ = np.zeros(ncol)
    dP = np.zeros(ncol)
    lib().lso_iterate_batch(C.byref(ctxs[0]._m), cols, ncol, int(niter), int(start_iter), _d(dJ), _d(dP))
    return dJ, dP


def max_threads():
    return lib().lso_max_threads()


def set_threads(n):
    """OpenMP thread count of iterate_batch (torchrun exports OMP_NUM_THREADS=1 for nproc > 1)."""
    lib().lso_set_threads(int(n))


def planck_bc(temperature, wavelength):
    """[Nspect, 2] table of planck(T[-2:], wav) (formal_solver.py:206) evaluated by the C restatement."""
    out = np.zeros((len(wavelength), 2))
    for la, wav in enumerate(wavelength):
        out[la, 0] = planck(temperature[-2], wav)
        out[la, 1] = planck(temperature[-1], wav)
    return out


__all__ = ['OracleContext', 'build', 'lib', 'w2', 'planck', 'piecewise_linear_1d', 'model_tables',
           'column_tables', 'iterate_batch', 'max_threads', 'set_threads', 'planck_bc', 'math']
