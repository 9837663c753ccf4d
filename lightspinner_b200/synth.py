"""Synthetic 1.5D column batches for benchmarks and scaling runs (BASELINE config 4 stand-in).

The GPU box has no Lightspinner reference, so the per-column set-up (witt EOS, background opacities, Voigt
profiles -- all outside the hot path) cannot be re-run there.  Instead a batch is derived from one
real reference column (a committed fixture, e.g. tests/golden/c2_falc_cah.npz = FALC, CaII + H active) by smooth,
column-specific, depth-dependent multiplicative jitter of the quantities that T / ne / vlos perturbations move:

    rng = default_rng(20260000 + col);  g1, g2, g3 = 5-point-boxcar-smoothed standard normals over depth
    fT = exp(0.01 g1), fne = exp(0.05 g2), fv = 1 + 0.02 g3
    bg_chi *= fne      bg_eta *= fne * fT      bg_sca *= fne      C *= fne
    phi[..., down, :] *= fv      phi[..., up, :] /= fv        (up/down asymmetry, as a non-zero vlos gives)

Every column therefore has its own background, collisional rates and line profiles (distinct memory, distinct
values) on the shared wavelength grid, starts from the fixture's populations and iterates like a real column.
The same recipe exists in two forms that produce bit-identical inputs: numpy on a problem dict (for the CPU
oracle) and torch on the device-resident host-pack blocks (for the GPU batch).
"""
import numpy as np


def jitter_factors(col, N):
    """[3, N] float64: fT, fne, fv of column `col` (column 0 is NOT special: every column is jittered)."""
    rng = np.random.default_rng(20260000 + int(col))

    def smooth(g):
        return np.convolve(np.pad(g, 2, 'edge'), np.ones(5) / 5, 'valid')

    g1, g2, g3 = (smooth(rng.standard_normal(N)) for _ in range(3))
    return np.stack([np.exp(0.01 * g1), np.exp(0.05 * g2), 1.0 + 0.02 * g3])


def jitter_atmosphere(col, temperature, ne):
    """BASELINE config 4's own recipe for synthetic column `col`, applied to the thermodynamic
    state as the reference applies it to its AtmosphereConstructor before convert_scales:
        T *= exp(0.01 g1),  ne *= exp(0.05 g2),  vlos = 2 km/s * g3        (hydrogen populations unchanged)
    with g1, g2, g3 = 5-point-boxcar-smoothed standard normals from default_rng(20260000 + col).
    Returns (T, ne, vlos [m/s])."""
    N = int(np.asarray(temperature).shape[0])
    rng = np.random.default_rng(20260000 + int(col))

    def smooth(g):
        return np.convolve(np.pad(g, 2, 'edge'), np.ones(5) / 5, 'valid')

    g1, g2, g3 = (smooth(rng.standard_normal(N)) for _ in range(3))
    return (np.asarray(temperature, dtype=np.float64) * np.exp(0.01 * g1),
            np.asarray(ne, dtype=np.float64) * np.exp(0.05 * g2), 2000.0 * g3)


def jitter_problem(p, col):
    """numpy form: returns a new problem dict for synthetic column `col` derived from base problem `p`."""
    N = int(p['Nspace'])
    fT, fne, fv = jitter_factors(col, N)
    q = dict(p)
    q['bg_chi'] = np.asarray(p['bg_chi']) * fne
    q['bg_eta'] = np.asarray(p['bg_eta']) * (fne * fT)
    q['bg_sca'] = np.asarray(p['bg_sca']) * fne
    q['C'] = np.asarray(p['C']) * fne
    phi = np.array(p['phi'], dtype=np.float64).reshape(-1, 2, N)
    phi[:, 0, :] = phi[:, 0, :] * fv
    phi[:, 1, :] = phi[:, 1, :] / fv
    q['phi'] = phi.reshape(-1)
    q['n'] = np.array(p['n'], dtype=np.float64, copy=True)
    return q


def jitter_staging(staging, lay, mt, base_pack, cols):
    """torch form: fills `staging` ([len(cols)][lay.hostpack] doubles on the device) with the host-pack blocks of
    the synthetic columns `cols`, derived from `base_pack` (one host-pack block on the device)."""
    import torch
    N, S = mt.Nspace, mt.Nspect
    n = len(cols)
    hp = lay.hostpack
    st = staging[:n * hp].view(n, hp)
    st.copy_(base_pack.view(1, hp).expand(n, hp))
    f = torch.from_numpy(np.stack([jitter_factors(c, N) for c in cols])).to(staging.device)   # [n, 3, N]
    fT, fne, fv = f[:, 0], f[:, 1], f[:, 2]

    def blk(off, rows):
        return st[:, off:off + rows * N].view(n, rows, N)

    blk(lay.hp_bg_chi, S).mul_(fne[:, None, :])
    blk(lay.hp_bg_eta, S).mul_((fne * fT)[:, None, :])
    blk(lay.hp_bg_sca, S).mul_(fne[:, None, :])
    blk(lay.hp_C, lay.sumNlevel2).mul_(fne[:, None, :])
    nphi = (lay.hp_wphi - lay.hp_phi) // (2 * N)
    phi = st[:, lay.hp_phi:lay.hp_phi + nphi * 2 * N].view(n, nphi, 2, N)
    phi[:, :, 0, :].mul_(fv[:, None, :])
    phi[:, :, 1, :].div_(fv[:, None, :])
    return st


def stress_problem(base, refine=10, nrays=10, ndepth=512):
    """BASELINE config 5 stand-in (the reference's grid generators are not on the GPU box): a single column derived
    from a real CaII/FALC problem `base` by
      * refining every wavelength interval `refine` times (transitions keep their index ranges, scaled),
      * replacing the angle quadrature by `nrays`-point Gauss-Legendre on [0, 1],
      * resampling all depth profiles to `ndepth` points, uniform in the base grid's index (log-interpolation for
        positive quantities, linear for T, height, velocities), vlos = 0,
      * recomputing the Voigt profiles phi and their normalisation wphi on the new grids (rh_method.py:198-243 with
        the stored damping parameters and Doppler widths).
    Both the CPU oracle and the CUDA path take the resulting arrays, so parity is checked on identical inputs."""
    from scipy import special
    from .tables import ModelTables, CLight
    S0, N0 = int(base['Nspect']), int(base['Nspace'])
    x0, x1 = np.arange(N0, dtype=np.float64), np.linspace(0.0, N0 - 1.0, ndepth)

    def depth(a, log):
        a = np.asarray(a, dtype=np.float64)
        flat = a.reshape(-1, N0)
        out = np.empty((flat.shape[0], ndepth))
        for r in range(flat.shape[0]):
            if log and np.all(flat[r] > 0):
                out[r] = np.exp(np.interp(x1, x0, np.log(flat[r])))
            else:
                out[r] = np.interp(x1, x0, flat[r])
        return out.reshape(a.shape[:-1] + (ndepth,))

    S1 = (S0 - 1) * refine + 1
    l0, l1 = np.arange(S0, dtype=np.float64), np.linspace(0.0, S0 - 1.0, S1)

    def lam(a, log):        # a[S0, ...] -> [S1, ...] along the first axis
        a = np.asarray(a, dtype=np.float64)
        flat = a.reshape(S0, -1)
        out = np.empty((S1, flat.shape[1]))
        for c in range(flat.shape[1]):
            col = flat[:, c]
            out[:, c] = np.exp(np.interp(l1, l0, np.log(col))) if (log and np.all(col > 0)) else np.interp(l1, l0, col)
        return out.reshape((S1,) + a.shape[1:])

    q = {k: base[k] for k in ('Nlevel', 'atom_names', 'linepar') if k in base}
    q['Nspace'], q['Nrays'], q['Nspect'] = int(ndepth), int(nrays), int(S1)
    xg, wg = np.polynomial.legendre.leggauss(nrays)
    q['muz'], q['wmu'] = 0.5 * (xg + 1.0), 0.5 * wg
    q['wavelength'] = np.interp(l1, l0, np.asarray(base['wavelength'], dtype=np.float64))
    q['height'] = depth(base['height'], False)
    q['temperature'] = depth(base['temperature'], False)
    for k in ('nStar', 'nTotal', 'n', 'C', 'hGround'):
        if k in base:
            q[k] = depth(base[k], True)
    vBroad = depth(base['vBroad'], False)
    aDamp = depth(base['aDamp'], True)
    q['vBroad'], q['aDamp'] = vBroad, aDamp
    q['vlos'] = np.zeros(ndepth)
    if 'vturb' in base:
        q['vturb'] = depth(base['vturb'], False)
    for k in ('bg_chi', 'bg_eta', 'bg_sca'):
        q[k] = np.ascontiguousarray(depth(lam(base[k], True), True))
    tr0 = np.asarray(base['trans'], dtype=np.int64).reshape(-1, 6)
    tr1 = tr0.copy()
    tr1[:, 4] = tr0[:, 4] * refine
    tr1[:, 5] = (tr0[:, 5] - 1) * refine + 1
    q['trans'] = tr1.astype(np.int32)
    off0 = np.concatenate([[0], np.cumsum(tr0[:, 5])])
    alpha = []
    for t in range(tr0.shape[0]):
        a0 = np.asarray(base['alpha'], dtype=np.float64)[off0[t]:off0[t + 1]]
        alpha.append(np.interp(np.linspace(0.0, len(a0) - 1.0, int(tr1[t, 5])), np.arange(len(a0), dtype=np.float64), a0))
    q['alpha'] = np.concatenate(alpha)
    q['phi'], q['phioff'], q['wphi'] = np.zeros(0), np.zeros(tr1.shape[0], dtype=np.int64), np.ones((tr1.shape[0], ndepth))
    mt = ModelTables(q)                       # wavelength weights on the refined grids
    phis, phioff, po = [], [], 0
    wphi = np.ones((tr1.shape[0], ndepth))
    sqrtPi = np.sqrt(np.pi)
    for t in range(tr1.shape[0]):
        atom, i, j, isLine, Nblue, Nlam = (int(v) for v in tr1[t])
        if not isLine:
            phioff.append(0)
            continue
        lambda0 = float(np.asarray(base['linepar']).reshape(-1, 4)[t, 3])
        wl = q['wavelength'][Nblue:Nblue + Nlam]
        vb = vBroad[atom]
        v = (wl[:, None] - lambda0) * CLight / (vb[None, :] * lambda0)
        ph = special.wofz(v + 1j * aDamp[t][None, :]).real / (sqrtPi * vb[None, :])        # [Nlam, ndepth]
        w = mt.wlambda[mt.toff[t]:mt.toff[t] + Nlam]
        wphi[t] = 1.0 / ((ph * w[:, None]).sum(axis=0) * q['wmu'].sum())
        full = np.broadcast_to(ph[:, None, None, :], (Nlam, nrays, 2, ndepth))
        phis.append(np.ascontiguousarray(full).reshape(-1))
        phioff.append(po)
        po += phis[-1].size
    q['phi'] = np.concatenate(phis)
    q['phioff'] = np.asarray(phioff, dtype=np.int64)
    q['wphi'] = wphi
    return q
