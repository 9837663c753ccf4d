#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Run here (the container that has /root/reference); the fixtures are committed so that the GPU box
(which has no reference) can check both the oracle and the CUDA path against real reference output.

    python tests/golden/make_golden.py [name ...]

Each fixture is an .npz holding (a) the flattened problem -- the exact arrays the reference's own
setup code produced (oracle/refharness/extract.py) -- and (b) what the reference's Context computed
from them: dJ/dPops history and J, I, Gamma, n at selected iterations of the test.py loop.

Fixtures
  c1_falc_ca        test.py's problem: CaII active, H passive, FALC, 5 rays (BASELINE configs 0/1); run to convergence
  c2_falc_cah       CaII + H active, FALC, 5 rays (C2 of SURVEY 8); run to convergence (79 iterations)
  c2v_jitter_cah_0, c2v_jitter_cah_1
                    BASELINE config 4's recipe, columns 0 and 1: T / ne jitter and non-zero vlos applied
                    before convert_scales, CaII + H active, 5 rays; run to convergence
  stress_r10_d512   BASELINE config 5's recipe, wavelength-reduced: every line's NlambdaGen x1 (the config says x10),
                    quadrature(10), FALC interpolated to 512 depths, CaII active; 8 iterations.  vlos = 0, so the line
                    profiles are stored once per (wavelength, depth) (`p_phi_compact`; helpers.load_golden expands them)
  c1v_jitter_ca3    CaII active, 3 rays, config-4 jitter recipe column 0 (T, ne, non-zero vlos): phi differs up/down
  rf_k40p, rf_k10m  response_fn.py columns: T[k] +/- 25 K, warm-started from the converged c1 populations
  setup_inputs      model-level atom data (levels, collisional-rate tables) and per-fixture ne / nHTot: inputs of the
                    device-side lte_pops / compute_collisions (their outputs nStar, C are in the column fixtures)
  eos               Wittmann EOS / background opacity / scale conversion known answers (FALC, a jitter column, a
                    response-function column) with the state of the witt() instance that produced them
  units             formal-solver / w2 / planck / uv known answers
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.refharness import load_reference, build_falc_setup  # noqa: E402
from oracle.refharness.extract import problem_from_reference_context  # noqa: E402


def jitter_modifier(col, ref):
    """SURVEY.md 8(d) item 4 jitter recipe."""
    u = sys.modules['astropy.units']

    def smooth(g):
        return np.convolve(np.pad(g, 2, 'edge'), np.ones(5) / 5, 'valid')

    def mod(ac):
        rng = np.random.default_rng(20260000 + col)
        n = ac.temperature.shape[0]
        g1, g2, g3 = (smooth(rng.standard_normal(n)) for _ in range(3))
        ac.temperature = (ac.temperature.value * np.exp(0.01 * g1)) << u.K
        ac.ne = (ac.ne.value * np.exp(0.05 * g2)) << (u.m ** -3)
        ac.vlos = (2000.0 * g3) << (u.m / u.s)
        return ac
    return mod


def run(ctx, max_iter, keep_full, keep_small):
    """test.py:20-29 loop with snapshots.  Returns dict of arrays."""
    out = {}
    hist = []
    dJ, dPops, i = 1.0, 1.0, 0
    while (dJ > 2e-3 or dPops > 1e-3) and i < max_iter:
        i += 1
        dJ = ctx.formal_sol_gamma_matrices()
        if i in keep_small or i in keep_full:
            # Gamma as left by formal_sol_gamma_matrices (before stat_equil touches n)
            out['it%d_Gamma' % i] = np.concatenate(
                [a.Gamma.reshape(a.Nlevel * a.Nlevel, -1) for a in ctx.activeAtoms], axis=0)
        if i > 3:
            dPops = ctx.stat_equil()
        hist.append((dJ, dPops))
        if i in keep_small or i in keep_full:
            out['it%d_I' % i] = np.array(ctx.I)
            out['it%d_n' % i] = np.concatenate([a.n for a in ctx.activeAtoms], axis=0)
            out['it%d_Jsum' % i] = np.array([ctx.J.sum(), np.abs(ctx.J).max()])
        if i in keep_full:
            out['it%d_J' % i] = np.array(ctx.J)
    out['hist'] = np.array(hist)
    out['niter'] = np.array(i)
    out['converged'] = np.array(not (dJ > 2e-3 or dPops > 1e-3))
    out['final_I'] = np.array(ctx.I)
    out['final_J'] = np.array(ctx.J)
    out['final_n'] = np.concatenate([a.n for a in ctx.activeAtoms], axis=0)
    return out


def save(name, problem, results, extra=None):
    d = {'p_' + k: v for k, v in problem.items()}
    d.update({'r_' + k: v for k, v in results.items()})
    if extra:
        d.update(extra)
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **d)
    print('%-18s %8.2f MB  niter=%s converged=%s' % (name, os.path.getsize(path) / 1e6, results.get('niter'),
                                                   results.get('converged')))


def make_c1(ref):
    atmos, spect, eqPops, bg = build_falc_setup(active=('Ca',), nrays=5)
    ctx = ref['rh_method'].Context(atmos, spect, eqPops, bg)
    p = problem_from_reference_context(ctx)
    r = run(ctx, 1000, keep_full={1, 4, 5}, keep_small={2, 3, 6, 10, 20, 30, 46})
    save('c1_falc_ca', p, r)
    return r['final_n']


def make_c2(ref):
    atmos, spect, eqPops, bg = build_falc_setup(active=('Ca', 'H'), nrays=5)
    ctx = ref['rh_method'].Context(atmos, spect, eqPops, bg)
    p = problem_from_reference_context(ctx)
    r = run(ctx, 1000, keep_full={1, 5}, keep_small={4, 6, 8, 12, 20, 40, 60, 79})
    save('c2_falc_cah', p, r)


def make_c2v(ref, col, with_phi):
    atmos, spect, eqPops, bg = build_falc_setup(active=('Ca', 'H'), nrays=5, modify_constructor=jitter_modifier(col, ref))
    ctx = ref['rh_method'].Context(atmos, spect, eqPops, bg)
    p = problem_from_reference_context(ctx)
    r = run(ctx, 1000, keep_full=set(), keep_small={1, 4, 6, 12, 20, 40})
    if not with_phi:
        # 4.9 MB of incompressible profiles: column 1 keeps what compute_phi consumes (aDamp, vBroad, vlos are in the
        # problem); tests re-form phi from them (helpers.recompute_phi = rh_method.py:198-243) or on the device
        p.pop('phi')
        p['phi_recompute'] = np.array(1)
    save('c2v_jitter_cah_%d' % col, p, r)


def make_stress(ref, refine=1, nrays=10, ndepth=512, niter=8):
    from oracle.refharness import falc_interpolated_constructor

    def refine_lines(models, ref):
        for m in models:
            for l in m.lines:
                l.NlambdaGen *= refine                       # SURVEY 8d item 5 (the full config uses 10)
            ref['atomic_model'].reconfigure_atom(m)          # atomic_model.py:71-72

    atmos, spect, eqPops, bg = build_falc_setup(active=('Ca',), nrays=nrays, modify_atoms=refine_lines,
                                                constructor=lambda: falc_interpolated_constructor(ndepth))
    ctx = ref['rh_method'].Context(atmos, spect, eqPops, bg)
    p = problem_from_reference_context(ctx)
    r = run(ctx, niter, keep_full={1}, keep_small={4, 5, 6, 8})
    del r['final_J']
    # vlos == 0: phi[la, mu, toFrom, k] does not depend on (mu, toFrom); keep one copy per (la, k)
    assert not np.any(p['vlos'])
    N, R = int(p['Nspace']), int(p['Nrays'])
    full = p.pop('phi').reshape(-1, R, 2, N)
    assert np.array_equal(full, np.broadcast_to(full[:, :1, :1, :], full.shape))
    p['phi_compact'] = np.ascontiguousarray(full[:, 0, 0, :])
    save('stress_r%d_d%d' % (nrays, ndepth), p, r)


def make_c1v(ref):
    atmos, spect, eqPops, bg = build_falc_setup(active=('Ca',), nrays=3, modify_constructor=jitter_modifier(0, ref))
    ctx = ref['rh_method'].Context(atmos, spect, eqPops, bg)
    p = problem_from_reference_context(ctx)
    r = run(ctx, 10, keep_full={1, 5}, keep_small={4, 6, 10})
    save('c1v_jitter_ca3', p, r)


def make_rf(ref, start_n, k, pert, name):
    u = sys.modules['astropy.units']

    def mod(ac):
        ac.temperature[k] += pert << u.K     # response_fn.py:26
        return ac
    atmos, spect, eqPops, bg = build_falc_setup(active=('Ca',), nrays=5, modify_constructor=mod,
                                                start_pops={'Ca': start_n})
    ctx = ref['rh_method'].Context(atmos, spect, eqPops, bg)
    p = problem_from_reference_context(ctx)
    r = run(ctx, 1000, keep_full={1}, keep_small={2, 4})
    save(name, p, r)


def make_setup_inputs(ref):
    """Inputs of the per-column set-up that precedes the hot path (lte_pops atomic_set.py:105-145,
    compute_collisions rh_method.py:474-487 / collisional_rates.py): the model-level data of the CaII and H6 atoms
    (levels, collisional-rate tables, abundances) and, for every column fixture, the electron and total hydrogen
    densities its atmosphere had (the fixtures already hold T, vturb and the reference's nStar, C, vBroad)."""
    from oracle.refharness import falc_interpolated_constructor
    out = {}
    models = {'CA': ref['rh_atoms'].CaII_atom(), 'H': ref['rh_atoms'].H_6_atom()}
    at = ref['atomic_set'].RadiativeSet(list(models.values())).atomicTable
    for nm, m in models.items():
        out['atom_%s_E_SI' % nm] = np.array([l.E_SI for l in m.levels])
        out['atom_%s_g' % nm] = np.array([l.g for l in m.levels])
        out['atom_%s_stage' % nm] = np.array([l.stage for l in m.levels], dtype=np.int32)
        out['atom_%s_abundance' % nm] = np.array(at[m.name].abundance)
        out['atom_%s_weight' % nm] = np.array(at[m.name].weight)
        kinds = {'Omega': 0, 'CI': 1, 'CE': 2}
        meta, T, R = [], [], []
        for c in m.collisions:
            meta.append([kinds[type(c).__name__], c.i, c.j, len(c.temperature)])
            T.append(np.asarray(c.temperature, dtype=np.float64))
            R.append(np.asarray(c.rates, dtype=np.float64))
        out['atom_%s_coll' % nm] = np.array(meta, dtype=np.int32)
        out['atom_%s_coll_T' % nm] = np.concatenate(T)
        out['atom_%s_coll_rates' % nm] = np.concatenate(R)

    def put(name, atmos):
        atmos.nondimensionalise()
        out[name + '_ne'] = np.array(atmos.ne)
        out[name + '_nHTot'] = np.array(atmos.nHTot)
        out[name + '_temperature'] = np.array(atmos.temperature)

    u = sys.modules['astropy.units']
    put('c1_falc_ca', build_falc_setup(active=('Ca',), nrays=5)[0])
    put('c2_falc_cah', build_falc_setup(active=('Ca', 'H'), nrays=5)[0])
    put('c1v_jitter_ca3', build_falc_setup(active=('Ca',), nrays=3, modify_constructor=jitter_modifier(0, ref))[0])
    for col in (0, 1):
        put('c2v_jitter_cah_%d' % col,
            build_falc_setup(active=('Ca', 'H'), nrays=5, modify_constructor=jitter_modifier(col, ref))[0])
    for k, pert, name in ((40, +25.0, 'rf_k40p'), (10, -25.0, 'rf_k10m')):
        def mod(ac, k=k, pert=pert):
            ac.temperature[k] += pert << u.K
            return ac
        put(name, build_falc_setup(active=('Ca',), nrays=5, modify_constructor=mod)[0])
    put('stress_r10_d512', build_falc_setup(active=('Ca',), nrays=10,
                                            constructor=lambda: falc_interpolated_constructor(512))[0])
    np.savez_compressed(os.path.join(HERE, 'setup_inputs.npz'), **out)
    print('setup_inputs       %8.2f MB' % (os.path.getsize(os.path.join(HERE, 'setup_inputs.npz')) / 1e6))


def make_eos(ref):
    """Known answers of the column set-up in front of everything else: the Wittmann EOS
    (witt.py:226-310, 342-432, 541-742), the background opacity (witt.py:744-1365 `cop`, background.py:21-53) and the
    column-mass -> height conversion (atmosphere.py:70-112), produced by ONE instrumented witt() instance whose state
    (abundances as normalised at that moment -- witt.__init__ renormalises the class-level table in place on every
    instantiation --, partition-function tables) is stored with its outputs."""
    import witt as W
    Const = ref['constants']
    at = ref['atomic_table'].get_global_atomic_table()
    eos = W.witt()
    out = {'ABUND': np.array(eos.ABUND), 'AMASS': np.array(eos.AMASS), 'avw': np.array(eos.avw), 'muH': np.array(eos.muH),
           'rho_from_H': np.array(eos.rho_from_H), 'ab_others': np.array(eos.ab_others), 'tpf': np.array(eos.tpf),
           'weightPerH': np.array(at.weightPerH)}
    nel = 28
    pf, eion, off = [], [], [0]
    for ii in range(nel):
        pf.append(np.array(eos.el[ii].pf))
        eion.append(np.array(eos.el[ii].eion))
        off.append(off[-1] + eos.el[ii].nstage)
    out['pf'] = np.concatenate(pf, axis=0)          # [sum nstage, npf]
    out['eion'] = np.concatenate(eion)
    out['stage_off'] = np.array(off, dtype=np.int32)
    u = sys.modules['astropy.units']
    cases = {'falc': None, 'jitter0': jitter_modifier(0, ref)}

    def rf(ac):
        ac.temperature[40] += 25.0 << u.K
        return ac
    cases['rf_k40p'] = rf
    wav = np.load(os.path.join(HERE, 'c2_falc_cah.npz'))['p_wavelength']
    out['wavelength'] = wav
    for name, mod in cases.items():
        ac = ref['fal'].Falc82()
        if mod is not None:
            ac = mod(ac) or ac
        ac.nondimensionalise()
        T, nHTot, ne, cmass = (np.array(x) for x in (ac.temperature, ac.nHTot, ac.ne, ac.depthScale))
        rho = Const.Amu * at.weightPerH * nHTot * Const.CM_TO_M**3 / Const.G_TO_KG        # background.py:33
        rhoSI = Const.Amu * at.weightPerH * nHTot                                             # atmosphere.py:82
        N = T.shape[0]
        pgas, pe = np.zeros(N), np.zeros(N)
        for k in range(N):
            pgas[k] = eos.pg_from_rho(T[k], rho[k])
            pe[k] = eos.pe_from_rho(T[k], rho[k])
        chi = np.zeros((wav.shape[0], N))
        chi_c = np.zeros(N)
        parts = np.zeros((N, 17))
        for k in range(N):
            chi[:, k] = eos.contOpacity(T[k], pgas[k], pe[k], wav * 10) / Const.CM_TO_M
            chi_c[k] = eos.contOpacity(T[k], pgas[k], pe[k], np.array([5000.0]))[0] / Const.CM_TO_M
            parts[k] = eos.getBackgroundPartials(T[k], pgas[k], pe[k], divide_by_u=True)
        eta = np.zeros_like(chi)
        for k in range(N):
            eta[:, k] = ref['utils'].planck(T[k], wav) * chi[:, k]
        height = np.zeros(N)
        tau_ref = np.zeros(N)
        tau_ref[0] = chi_c[0] / rhoSI[0] * cmass[0]                                          # atmosphere.py:98-106
        for k in range(1, N):
            height[k] = height[k - 1] - 2.0 * (cmass[k] - cmass[k - 1]) / (rhoSI[k - 1] + rhoSI[k])
            tau_ref[k] = tau_ref[k - 1] + 0.5 * (chi_c[k - 1] + chi_c[k]) * (height[k - 1] - height[k])
        hTau1 = np.interp(1.0, tau_ref, height)
        height -= hTau1
        class A:
            pass
        a = A()
        a.ne = ne
        sca1 = ref['background'].thomson_scattering(a)
        for k, v in (('T', T), ('nHTot', nHTot), ('ne', ne), ('cmass', cmass), ('rho', rho), ('pgas', pgas), ('pe', pe),
                     ('chi', chi), ('eta', eta), ('thomson', sca1), ('chi_c', chi_c), ('partials', parts),
                     ('height', height), ('tau_ref', tau_ref)):
            out['%s_%s' % (name, k)] = v
    np.savez_compressed(os.path.join(HERE, 'eos.npz'), **out)
    print('eos                %8.2f MB' % (os.path.getsize(os.path.join(HERE, 'eos.npz')) / 1e6))


def make_units(ref):
    fs = ref['formal_solver']
    rng = np.random.default_rng(12345)
    out = {}
    # w2 at its three branches (formal_solver.py:34-43)
    dt = np.array([1e-9, 1e-6, 4.9999e-4, 5e-4, 5.0001e-4, 1e-3, 0.1, 1.0, 7.3, 49.9, 50.0, 50.0001, 80.0, 1e4])
    out['w2_dtau'] = dt
    out['w2_w'] = np.array([fs.w2(x) for x in dt])
    # planck
    T = np.array([4170.0, 5000.0, 6520.0, 9400.0, 1.0e5])
    wav = np.array([30.0, 91.17, 121.567, 393.366, 500.0, 854.209, 2000.0])
    out['planck_T'] = T
    out['planck_wav'] = wav
    out['planck_B'] = np.array([[ref['utils'].planck(np.array([t]), w)[0] for w in wav] for t in T])

    # piecewise_linear_1d on synthetic atmospheres covering all w2 branches
    class A:
        pass
    cases = []
    for N, scale in ((82, 1e-7), (82, 1e-4), (33, 1e-2), (7, 1.0), (3, 1e-6), (200, 1e-5)):
        a = A()
        a.Nspace = N
        a.height = np.sort(rng.uniform(0, 2.0e6, N))[::-1].copy()
        a.temperature = rng.uniform(4000, 9000, N)
        a.muz = np.array([0.0469100770306680, 0.5, 0.9530899229693319])
        chi = scale * np.exp(rng.normal(0, 2.0, N)) * np.exp(np.linspace(-6, 6, N))
        S = rng.uniform(1e-9, 5e-8, N)
        for mu in range(3):
            for toFrom in (0, 1):
                wv = float(rng.uniform(100, 900))
                r = fs.piecewise_linear_1d(a, mu, toFrom, wv, chi, S)
                cases.append((a.height, a.temperature, a.muz[mu], toFrom, wv, chi, S, r.I, r.PsiStar))
    out['fs_ncase'] = np.array(len(cases))
    for i, c in enumerate(cases):
        for nm, v in zip(('z', 'T', 'mu', 'toFrom', 'wav', 'chi', 'S', 'I', 'Psi'), c):
            out['fs%d_%s' % (i, nm)] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, 'units.npz'), **out)
    print('units              %8.2f MB' % (os.path.getsize(os.path.join(HERE, 'units.npz')) / 1e6))


if __name__ == '__main__':
    which = set(sys.argv[1:])
    ref = load_reference()
    t0 = time.time()
    if not which or 'units' in which:
        make_units(ref)
    final_n = None
    if not which or which & {'c1', 'rf'}:
        final_n = make_c1(ref)
    if not which or 'rf' in which:
        make_rf(ref, final_n, 40, +25.0, 'rf_k40p')
        make_rf(ref, final_n, 10, -25.0, 'rf_k10m')
    if not which or 'c2' in which:
        make_c2(ref)
    if not which or 'c1v' in which:
        make_c1v(ref)
    if not which or 'c2v' in which:
        make_c2v(ref, 0, True)
        make_c2v(ref, 1, False)
    if not which or 'stress' in which:
        make_stress(ref)
    if not which or 'setup' in which:
        make_setup_inputs(ref)
    if not which or 'eos' in which:
        make_eos(ref)
    print('done in %.0f s' % (time.time() - t0))
