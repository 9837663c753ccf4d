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
    if 'phi_compact' in p:      # vlos == 0 fixtures keep one profile per (wavelength, depth): expand to the reference's shape
        c = p.pop('phi_compact')
        p['phi'] = np.ascontiguousarray(np.broadcast_to(c[:, None, None, :], (c.shape[0], p['Nrays'], 2, c.shape[1]))).reshape(-1)
    if 'phi_recompute' in p:    # fixture without the (incompressible) profiles: re-form them as the reference does
        p.pop('phi_recompute')
        p['phi'], p['wphi'] = recompute_phi(p)
    return p, r


def recompute_phi(p):
    """ComputationalTransition.compute_phi (rh_method.py:198-243) from a problem's aDamp / vBroad / vlos with the
    reference's own expressions (scipy wofz): (phi concat [Nlam, Nrays, 2, N] per line, wphi [Ntrans, N])."""
    from scipy import special
    CLight = 2.99792458E+08
    N, R = int(p['Nspace']), int(p['Nrays'])
    trans = np.asarray(p['trans']).reshape(-1, 6)
    wav, muz, wmu = np.asarray(p['wavelength']), np.asarray(p['muz']), np.asarray(p['wmu'])
    sqrtPi = np.sqrt(np.pi)
    phis, wphi = [], np.zeros((trans.shape[0], N))
    for t, (atom, i, j, isLine, Nblue, Nlam) in enumerate(trans):
        if not isLine:
            continue
        lambda0 = np.asarray(p['linepar']).reshape(-1, 4)[t, 3]
        vBroad, aDamp = np.asarray(p['vBroad'])[atom], np.asarray(p['aDamp'])[t]
        wl = wav[Nblue:Nblue + Nlam]
        w = np.zeros(Nlam)
        w[0], w[-1], w[1:-1] = 0.5 * (wl[1] - wl[0]), 0.5 * (wl[-1] - wl[-2]), 0.5 * (wl[2:] - wl[:-2])
        w = (CLight / lambda0) * w
        phi = np.zeros((Nlam, R, 2, N))
        wPhi = np.zeros(N)
        vlosDop = np.stack([muz[mu] * np.asarray(p['vlos']) / vBroad for mu in range(R)])
        for la in range(Nlam):
            v = (wl[la] - lambda0) * CLight / (vBroad * lambda0)
            for mu in range(R):
                wlamu = w * 0.5 * wmu[mu]
                for toFrom, sign in enumerate([-1.0, 1.0]):
                    vk = v + sign * vlosDop[mu]
                    phi[la, mu, toFrom, :] = special.wofz(vk + 1j * aDamp).real / (sqrtPi * vBroad)
                    wPhi[:] += phi[la, mu, toFrom, :] * wlamu[la]
        phis.append(phi.reshape(-1))
        wphi[t] = 1.0 / wPhi
    return (np.concatenate(phis) if phis else np.zeros(0)), wphi


def load_setup_inputs(name):
    """(atoms, column): model-level data of the CaII / H6 atoms {'CA': {...}, 'H': {...}} and the ne, nHTot, temperature
    of fixture `name`'s atmosphere (tests/golden/setup_inputs.npz)."""
    z = np.load(os.path.join(GOLDEN, 'setup_inputs.npz'))
    atoms = {}
    for nm in ('CA', 'H'):
        pre = 'atom_%s_' % nm
        atoms[nm] = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        atoms[nm]['abundance'] = float(atoms[nm]['abundance'])
        atoms[nm]['weight'] = float(atoms[nm]['weight'])
    col = {k: z['%s_%s' % (name, k)] for k in ('ne', 'nHTot', 'temperature')}
    return atoms, col


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


# ----------------------------------------------------------------------------------------------------------
# Fixture-backed stand-ins for the reference's model objects (atmosphere, atoms, transitions, populations), so that
# the drop-in lightspinner_b200.Context can be driven exactly like rh_method.Context on a box that has no reference.
class _FakeAtmos:
    def __init__(self, p):
        self.height = np.array(p['height'])
        self.temperature = np.array(p['temperature'])
        self.vlos = np.array(p['vlos'])
        self.vturb = np.array(p['vturb'])
        self.muz = np.array(p['muz'])
        self.wmu = np.array(p['wmu'])
        self.dimensioned = True

    @property
    def Nspace(self):
        return self.height.shape[0]

    @property
    def Nrays(self):
        return self.muz.shape[0]

    def nondimensionalise(self):
        self.dimensioned = False

    def dimensionalise(self):
        self.dimensioned = True


class _FakeLine:
    def __init__(self, i, j, par, wavelength, aDamp):
        self.i, self.j = int(i), int(j)
        self.Aji, self.Bji, self.Bij, self.lambda0 = (np.float64(v) for v in par)
        self.wavelength = wavelength
        self._aDamp = aDamp

    def damping(self, atmos, vBroad, hGround):
        return self._aDamp, None


class _FakeContinuum:
    def __init__(self, i, j, alpha, wavelength):
        self.i, self.j = int(i), int(j)
        self.alpha = alpha
        self.wavelength = wavelength


class _FakeCollisions:
    def __init__(self, C):
        self._C = C

    def compute_rates(self, atmos, nStar, C):
        C += self._C


class _FakeAtom:
    def __init__(self, name, Nlevel, vBroad, C):
        self.name = name
        self.levels = [None] * int(Nlevel)
        self.lines, self.continua = [], []
        self.collisions = [_FakeCollisions(C)]
        self._vBroad = vBroad

    def v_broad(self, atmos):
        return self._vBroad


class _FakeState:
    def __init__(self, nStar, nTotal, pops=None):
        self.nStar, self.nTotal, self.pops = nStar, nTotal, pops

    @property
    def n(self):
        return self.pops if self.pops is not None else self.nStar


class _FakePops(dict):
    atomicTable = None

    def __getitem__(self, name):
        return dict.__getitem__(self, name.upper().strip())


class _NS:
    pass


def fake_reference_objects(p):
    """(atmos, spect, eqPops, background) built from a fixture problem dict: duck-typed like the reference's."""
    N = int(p['Nspace'])
    atmos = _FakeAtmos(p)
    names = [str(s).strip() for s in p['atom_names']]
    lvloff = np.concatenate([[0], np.cumsum(p['Nlevel'])]).astype(int)
    g2off = np.concatenate([[0], np.cumsum(np.asarray(p['Nlevel'], dtype=int) ** 2)]).astype(int)
    toff = np.concatenate([[0], np.cumsum(p['trans'][:, 5])]).astype(int)
    atoms = []
    for a, nm in enumerate(names):
        NL = int(p['Nlevel'][a])
        C = np.array(p['C'][g2off[a]:g2off[a + 1]]).reshape(NL, NL, N)
        atoms.append(_FakeAtom(nm, NL, np.array(p['vBroad'][a]), C))
    transitions = []
    wavelength = np.array(p['wavelength'])
    for t, (atom, i, j, isLine, Nblue, Nlam) in enumerate(p['trans']):
        wl = np.copy(wavelength[Nblue:Nblue + Nlam])
        if isLine:
            tr = _FakeLine(i, j, p['linepar'][t], wl, np.array(p['aDamp'][t]))
            atoms[atom].lines.append(tr)
        else:
            tr = _FakeContinuum(i, j, np.array(p['alpha'][toff[t]:toff[t + 1]]), wl)
            atoms[atom].continua.append(tr)
        tr._range = (int(Nblue), int(Nblue + Nlam))
        transitions.append(tr)
    spect = _NS()
    spect.wavelength = wavelength
    spect.transitions = transitions
    spect.activeSet = [[tr for tr in transitions if tr._range[0] <= la < tr._range[1]] for la in range(len(wavelength))]
    spect.radSet = _NS()
    spect.radSet.activeAtoms = atoms
    eqPops = _FakePops()
    for a, nm in enumerate(names):
        nStar = np.array(p['nStar'][lvloff[a]:lvloff[a + 1]])
        n0 = np.array(p['n'][lvloff[a]:lvloff[a + 1]])
        eqPops[nm] = _FakeState(nStar, np.array(p['nTotal'][a]), None if np.array_equal(n0, nStar) else n0)
    if 'H' not in eqPops:
        hn = np.zeros((1, N))
        hn[0] = p['hGround']
        eqPops['H'] = _FakeState(hn, hn[0].copy(), None)
    bg = _NS()
    bg.chi, bg.eta, bg.sca = np.array(p['bg_chi']), np.array(p['bg_eta']), np.array(p['bg_sca'])
    return atmos, spect, eqPops, bg


# ----------------------------------------------------------------------------------------------------------
# Mutated fixtures: ragged / edge-case problems derived from a reference fixture.  They are not outputs of the
# reference any more, so the checker for them is the oracle (which is pinned to the reference on the originals).
def drop_depth(p, k):
    """The same problem with depth point k removed (odd Nspace -> no TMA staging path, different tile tails)."""
    N = int(p['Nspace'])
    keep = np.array([q for q in range(N) if q != k])
    q = dict(p)
    q['Nspace'] = N - 1
    for key in ('height', 'temperature', 'vlos', 'vturb'):
        if key in p:
            q[key] = np.array(p[key])[keep]
    for key in ('bg_chi', 'bg_eta', 'bg_sca', 'nStar', 'nTotal', 'C', 'n', 'wphi', 'aDamp', 'vBroad'):
        if key in p:
            q[key] = np.ascontiguousarray(np.array(p[key])[..., keep])
    phi = np.array(p['phi']).reshape(-1, N)[:, keep]
    q['phi'] = np.ascontiguousarray(phi).reshape(-1)
    q['phioff'] = (np.array(p['phioff'], dtype=np.int64) // N) * (N - 1)
    return q


def select_rays(p, rays):
    """The same problem restricted to the given angle indices (e.g. a single ray: 32 wavelengths per warp)."""
    N, R = int(p['Nspace']), int(p['Nrays'])
    rays = list(rays)
    q = dict(p)
    q['Nrays'] = len(rays)
    q['muz'] = np.array(p['muz'])[rays]
    q['wmu'] = np.array(p['wmu'])[rays]
    phi = np.array(p['phi']).reshape(-1, R, 2, N)[:, rays]
    q['phi'] = np.ascontiguousarray(phi).reshape(-1)
    q['phioff'] = (np.array(p['phioff'], dtype=np.int64) // R) * len(rays)
    return q


def only_transitions(p, keep):
    """The same problem with only the listed transitions radiatively active (possibly none)."""
    keep = list(keep)
    tr = np.array(p['trans']).reshape(-1, 6)
    N, R = int(p['Nspace']), int(p['Nrays'])
    toff = np.concatenate([[0], np.cumsum(tr[:, 5])]).astype(int)
    q = dict(p)
    q['trans'] = tr[keep].reshape(-1, 6)
    q['linepar'] = np.array(p['linepar']).reshape(-1, 4)[keep].reshape(-1, 4)
    q['alpha'] = np.concatenate([np.array(p['alpha'])[toff[t]:toff[t + 1]] for t in keep]) if keep else np.zeros(0)
    q['wphi'] = np.array(p['wphi'])[keep].reshape(-1, N)
    phis, offs, o = [], [], 0
    for t in keep:
        if tr[t, 3]:
            sz = int(tr[t, 5]) * R * 2 * N
            po = int(p['phioff'][t])
            phis.append(np.array(p['phi'])[po:po + sz])
            offs.append(o)
            o += sz
        else:
            offs.append(0)
    q['phi'] = np.concatenate(phis) if phis else np.zeros(0)
    q['phioff'] = np.array(offs, dtype=np.int64)
    return q
