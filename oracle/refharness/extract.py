"""Flatten a live reference `rh_method.Context` into the plain-dict problem format of
oracle/mali_oracle.py (TEST INFRASTRUCTURE ONLY; used by tests/golden/make_golden.py and by the
CPU tests that run next to /root/reference)."""
import numpy as np


def problem_from_reference_context(ctx, copy_pops=True):
    atmos, spect, bg = ctx.atmos, ctx.spect, ctx.background
    N = atmos.Nspace
    Nlevel, trans, linepar, alpha, phi, phioff, wphi = [], [], [], [], [], [], []
    nStar, nTotal, Cs, ns = [], [], [], []
    aDamp, vBroad = [], []
    off = 0
    for ia, atom in enumerate(ctx.activeAtoms):
        Nlevel.append(atom.Nlevel)
        atom.compute_collisions()          # rh_method.py:474-487 (iteration-invariant)
        nStar.append(np.array(atom.nStar))
        nTotal.append(np.array(atom.nTotal))
        Cs.append(np.array(atom.C).reshape(atom.Nlevel * atom.Nlevel, N))
        ns.append(np.array(atom.n, copy=copy_pops))
        vBroad.append(np.array(atom.vBroad))
        for t in atom.trans:
            Nlam = t.wavelength.shape[0]
            assert np.array_equal(t.wavelength, spect.wavelength[t.Nblue:t.Nblue + Nlam])
            act = np.zeros(spect.wavelength.shape[0], bool)
            act[t.Nblue:t.Nblue + Nlam] = True
            assert np.array_equal(act, t.active), 'transition not active on a contiguous range'
            trans.append([ia, t.i, t.j, int(t.isLine), int(t.Nblue), Nlam])
            if t.isLine:
                linepar.append([t.Aji, t.Bji, t.Bij, t.lambda0])
                alpha.append(np.zeros(Nlam))
                phioff.append(off)
                phi.append(np.ascontiguousarray(t.phi).ravel())
                off += t.phi.size
                wphi.append(np.array(t.wphi))
                aDamp.append(np.array(t.transModel.damping(atmos, atom.vBroad, atom.hPops.n[0])[0]))
            else:
                aDamp.append(np.zeros(N))
                linepar.append([0.0, 0.0, 0.0, 0.0])
                alpha.append(np.array(t.alpha, dtype=np.float64))
                phioff.append(0)
                wphi.append(np.zeros(N))
    return dict(
        Nspace=N, Nrays=atmos.Nrays, Nspect=spect.wavelength.shape[0],
        wavelength=np.array(spect.wavelength), muz=np.array(atmos.muz), wmu=np.array(atmos.wmu),
        atom_names=np.array([a.atomicModel.name for a in ctx.activeAtoms]),
        Nlevel=np.array(Nlevel, dtype=np.int32), trans=np.array(trans, dtype=np.int32),
        linepar=np.array(linepar), alpha=np.concatenate(alpha),
        height=np.array(atmos.height), temperature=np.array(atmos.temperature),
        vlos=np.array(atmos.vlos), vturb=np.array(atmos.vturb), aDamp=np.stack(aDamp), vBroad=np.stack(vBroad),
        hGround=np.array(ctx.activeAtoms[0].hPops.n[0]),
        bg_chi=np.array(bg.chi), bg_eta=np.array(bg.eta), bg_sca=np.array(bg.sca),
        nStar=np.concatenate(nStar, axis=0), nTotal=np.stack(nTotal), C=np.concatenate(Cs, axis=0),
        n=np.concatenate(ns, axis=0),
        phi=np.concatenate(phi) if phi else np.zeros(0), phioff=np.array(phioff, dtype=np.int64),
        wphi=np.stack(wphi))
