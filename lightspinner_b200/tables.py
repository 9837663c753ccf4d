"""Host-side flattening of one MALI problem into the arrays libmali_b200.so consumes.

A *problem* is a dict of numpy arrays in the reference's own (depth-contiguous) layouts -- the state a
reference `Context` holds after construction:

model part (shared by every column of a batch)
    Nspace, Nrays, Nspect      ints
    wavelength [Nspect] (nm), muz [Nrays], wmu [Nrays]
    Nlevel [Natom]             levels per active atom, in Context.activeAtoms order        (rh_method.py:558-560)
    trans  [Ntrans, 6]         atom, i, j, isLine, Nblue, Nlambda -- atom.trans order       (rh_method.py:398-405)
    linepar [Ntrans, 4]        Aji, Bji, Bij, lambda0 (zeros for continua)
    alpha  [sum Nlambda]       continuum cross-sections, concatenated over transitions (zeros for lines)
column part
    height, temperature [Nspace]
    bg_chi, bg_eta, bg_sca [Nspect, Nspace]                                                 (background.py:37-51)
    nStar [sumNlevel, Nspace], nTotal [Natom, Nspace], n [sumNlevel, Nspace]
    C [sum Nlevel^2, Nspace]   C[i, j, k] per atom, flattened                               (rh_method.py:474-487)
    phi  1-D concat of each line's phi[Nlambda, Nrays, 2, Nspace], phioff [Ntrans]           (rh_method.py:224)
    wphi [Ntrans, Nspace]      (rows of continua unused)
optional
    wlambda_table [sum Nlambda]  the wavelength weights, when a transition's grid here is only a slice of the
                                 grid they were formed on (wavelength-sharded single column, lambda_shard.py)

Everything here is iteration-invariant set-up.  The tables that the reference evaluates with numpy
transcendental code (not numba) are evaluated here with the same scalar/vector numpy expressions so that
they carry the same bits; the CUDA kernels then only do +, -, *, / and exp(-dtau).
"""
import ctypes as C

import numpy as np

from . import _capi

# constants.py:1-4,17
CLight = 2.99792458E+08
HPlanck = 6.6260755E-34
HC = HPlanck * CLight
KBoltzmann = 1.380658E-23
NM_TO_M = 1.0E-09


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def transition_offsets(trans):
    off = np.zeros(trans.shape[0] + 1, dtype=np.int64)
    np.cumsum(trans[:, 5], out=off[1:])
    return off


class ModelTables:
    """Per-model constants: uv prefactors (rh_method.py:268-286), wavelength weights (:157-196, :451, :455)."""

    def __init__(self, p):
        self.Nspace = int(p['Nspace'])
        self.Nrays = int(p['Nrays'])
        self.Nspect = int(p['Nspect'])
        self.Nlevel = np.ascontiguousarray(p['Nlevel'], dtype=np.int32)
        self.Natom = int(self.Nlevel.shape[0])
        self.trans = np.ascontiguousarray(p['trans'], dtype=np.int32).reshape(-1, 6)
        self.Ntrans = int(self.trans.shape[0])
        self.wavelength = _f64(p['wavelength'])
        self.muz = _f64(p['muz'])
        self.wmu = _f64(p['wmu'])
        self.linepar = _f64(p['linepar']).reshape(-1, 4)
        self.alpha = _f64(p['alpha'])
        self.toff = transition_offsets(self.trans)
        self.lvloff = np.concatenate([[0], np.cumsum(self.Nlevel)]).astype(np.int64)
        self.g2off = np.concatenate([[0], np.cumsum(self.Nlevel.astype(np.int64)**2)]).astype(np.int64)
        ntab = int(self.toff[-1])
        if self.alpha.shape[0] != ntab:
            raise ValueError('alpha must hold sum(Nlambda) = %d entries' % ntab)
        self.lineconst = np.zeros((self.Ntrans, 3))
        self.lambda0 = np.ascontiguousarray(self.linepar[:, 3])      # line-centre wavelengths (device compute_phi)
        self.wlambda = np.zeros(ntab)
        self.twohc_l3 = np.zeros(ntab)
        self.wlacont = np.zeros(ntab)
        hc_4pi = 0.25 * HC / np.pi                                   # rh_method.py:268
        w_given = p.get('wlambda_table')
        if w_given is not None:
            w_given = _f64(w_given)
            if w_given.shape[0] != ntab:
                raise ValueError('wlambda_table must hold sum(Nlambda) = %d entries' % ntab)
        for t in range(self.Ntrans):
            atom, i, j, isLine, Nblue, Nlam = (int(v) for v in self.trans[t])
            wl = self.wavelength[Nblue:Nblue + Nlam]
            o = int(self.toff[t])
            if isLine:
                Aji, Bji, Bij, lambda0 = (np.float64(v) for v in self.linepar[t])
                self.lineconst[t, 0] = hc_4pi * Bij                  # :279
                self.lineconst[t, 1] = Aji / Bji                     # :281
                self.lineconst[t, 2] = Bji / Bij                     # :450
                dopplerWidth = CLight / lambda0                      # :179
            else:
                dopplerWidth = 1.0
            if w_given is not None:
                w = w_given[o:o + Nlam]
            else:
                if Nlam < 2:
                    raise ValueError('transition %d has %d wavelength(s): its quadrature weights need at least 2 '
                                     '(a wavelength shard that clips a transition must pass wlambda_table)' % (t, Nlam))
                w = np.empty(Nlam)
                w[0] = 0.5 * (wl[1] - wl[0]) * dopplerWidth          # :185
                w[-1] = 0.5 * (wl[-1] - wl[-2]) * dopplerWidth       # :187
                w[1:-1] = 0.5 * (wl[2:] - wl[:-2]) * dopplerWidth    # :189
            self.wlambda[o:o + Nlam] = w
            if not isLine:
                for lt in range(Nlam):                               # scalar numpy power / divide, as the reference
                    self.twohc_l3[o + lt] = 2.0 * HC / (NM_TO_M * wl[lt])**3     # :286
                    self.wlacont[o + lt] = w[lt] / wl[lt] / HPlanck              # :455

    def desc(self):
        """ctypes mali_model_desc viewing this object's arrays (keep `self` alive while it is used)."""
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        return _capi.ModelDesc(self.Nspace, self.Nrays, self.Nspect, self.Natom, self.Ntrans, ip(self.Nlevel),
                               ip(self.trans), dp(self.wavelength), dp(self.muz), dp(self.wmu), dp(self.lineconst),
                               dp(self.wlambda), dp(self.alpha), dp(self.twohc_l3), dp(self.wlacont), dp(self.lambda0))


def gij_continuum(mt, nStar, temperature):
    """g_ij of every continuum, [sum over continua of Nlambda, Nspace], exactly as rh_method.py:453-454."""
    hc_k = HC / (KBoltzmann * NM_TO_M)
    rows = []
    temperature = _f64(temperature)
    nStar = _f64(nStar)
    for t in range(mt.Ntrans):
        atom, i, j, isLine, Nblue, Nlam = (int(v) for v in mt.trans[t])
        if isLine:
            continue
        ni = np.ascontiguousarray(nStar[mt.lvloff[atom] + i])
        nj = np.ascontiguousarray(nStar[mt.lvloff[atom] + j])
        g = np.empty((Nlam, mt.Nspace))
        for lt in range(Nlam):
            g[lt] = ni / nj * np.exp(-hc_k / mt.wavelength[Nblue + lt] / temperature)
        rows.append(g)
    if not rows:
        return np.zeros((0, mt.Nspace))
    return np.concatenate(rows, axis=0)


def planck_bc(wavelength, temperature):
    """[Nspect, 2]: planck(T[-2:], wav) of formal_solver.py:206 (utils.py:17-22 as numba compiles it: y*y*y, libm exp)."""
    wavelength = _f64(wavelength)
    out = np.empty((wavelength.shape[0], 2))
    dp = C.POINTER(C.c_double)
    _capi.check(_capi.load().mali_planck_bc(wavelength.ctypes.data_as(dp), wavelength.shape[0],
                                            float(temperature[-2]), float(temperature[-1]),
                                            out.ctypes.data_as(dp)))
    return out


def pack_column(mt, lay, p, out=None, with_phi=True, with_derived=True, with_background=True):
    """Concatenate one column's reference-layout arrays into the host staging block (`mali_layout.hp_*`).

    No transposition happens on the host: mali_upload_columns re-lays the data out on the device.
    with_phi=False: only the first `lay.hp_phi` doubles (everything but the line profiles phi / wphi, which
    mali_compute_phi then forms on the device).
    with_derived=False: only the first `lay.hp_C` doubles -- heights, boundary Planck values, background, nTotal --
    (mali_setup_columns forms C, the continua's g_ij and the LTE populations on the device as well).
    with_background=False: only the first `lay.hp_bg_chi` doubles -- heights (zeros if the problem has none: a
    column-mass scale that mali_background converts), boundary Planck values, nTotal.
    """
    N, Nspect = mt.Nspace, mt.Nspect
    if not with_background:
        with_derived = False
    if not with_derived:
        with_phi = False
    size = lay.hostpack if with_phi else (lay.hp_phi if with_derived else (lay.hp_C if with_background else lay.hp_bg_chi))
    if out is None:
        out = np.empty(size)
    if out.shape[0] != size:
        raise ValueError('host pack must hold %d doubles' % size)

    def put(off, arr, size):
        a = np.asarray(arr, dtype=np.float64)
        if a.size != size:
            raise ValueError('array of %d elements where %d were expected' % (a.size, size))
        out[off:off + size] = a.reshape(-1)

    put(lay.hp_height, p['height'] if 'height' in p else np.zeros(N), N)
    put(lay.hp_bbc, planck_bc(mt.wavelength, _f64(p['temperature'])), 2 * Nspect)
    put(lay.hp_nTotal, p['nTotal'], mt.Natom * N)
    if not with_background:
        return out
    put(lay.hp_bg_chi, p['bg_chi'], Nspect * N)
    put(lay.hp_bg_eta, p['bg_eta'], Nspect * N)
    put(lay.hp_bg_sca, p['bg_sca'], Nspect * N)
    if not with_derived:
        return out
    put(lay.hp_C, p['C'], lay.sumNlevel2 * N)
    if with_phi:
        phi = _f64(p['phi'])
        o = lay.hp_phi
        for t in range(mt.Ntrans):
            atom, i, j, isLine, Nblue, Nlam = (int(v) for v in mt.trans[t])
            if isLine:
                sz = Nlam * mt.Nrays * 2 * N
                po = int(p['phioff'][t])
                out[o:o + sz] = phi[po:po + sz]
                o += sz
        put(lay.hp_wphi, p['wphi'], mt.Ntrans * N)
    g = gij_continuum(mt, p['nStar'], p['temperature'])
    put(lay.hp_gijcont, g, g.size)
    if lay.hp_gijcont + g.size != lay.hp_n:
        raise ValueError('continuum table size mismatch')
    put(lay.hp_n, p['n'], lay.sumNlevel * N)
    return out
