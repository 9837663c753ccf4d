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
