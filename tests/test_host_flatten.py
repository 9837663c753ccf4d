"""Host-side logic of the drop-in Context (no GPU): the mirror of the reference's per-Context set-up (damping
parameters and Doppler widths asked of the model objects, collisions rh_method.py:474-487, activity ranges :122-127)
and the flattening / host-pack layout must reproduce what the reference computed, bit for bit.  The line profiles
themselves (rh_method.py:198-243) are formed on the device; tests/test_gpu_device_phi.py checks them."""
import numpy as np
import pytest

from helpers import fake_reference_objects, load_golden, select_rays

KEYS = ['wavelength', 'muz', 'wmu', 'Nlevel', 'trans', 'linepar', 'alpha', 'height', 'temperature', 'bg_chi', 'bg_eta',
        'bg_sca', 'nStar', 'nTotal', 'C', 'n', 'aDamp', 'vBroad', 'vlos']


@pytest.mark.parametrize('name', ['c1_falc_ca', 'c2_falc_cah', 'c1v_jitter_ca3', 'rf_k40p'])
def test_context_flattening_reproduces_reference_arrays(name):
    """Drive lightspinner_b200.Context's host mirror with fixture-backed stand-ins of the reference objects."""
    from lightspinner_b200.context import Context
    p, r = load_golden(name)
    atmos, spect, eqPops, bg = fake_reference_objects(p)
    ctx = Context(atmos, spect, eqPops, bg, _host_only=True)
    q = ctx._problem
    for k in KEYS:
        assert np.array_equal(np.asarray(q[k]), np.asarray(p[k])), (name, k)
    # aliasing contract of rh_method.py:411-416
    for atom in ctx.activeAtoms:
        assert eqPops[atom.atomicModel.name].pops is atom.n
    with pytest.raises(Exception):
        ctx.formal_sol_gamma_matrices()      # no engine: the host-only object cannot compute (no CPU fallback)


def test_model_tables_match_oracle_tables(oracle):
    from lightspinner_b200 import tables
    for name in ['c1_falc_ca', 'c2_falc_cah']:
        p, r = load_golden(name)
        mt = tables.ModelTables(p)
        ot = oracle.model_tables(p)
        assert np.array_equal(mt.lineconst, ot['lineconst'])
        assert np.array_equal(mt.wlambda, ot['wlambda'])
        assert np.array_equal(mt.twohc_l3, ot['twohc_l3'])
        assert np.array_equal(mt.wlacont, ot['wlacont'])
        g = tables.gij_continuum(mt, p['nStar'], p['temperature'])
        og = oracle.column_tables(p)['gijcont']
        rows = np.concatenate([np.arange(ot['toff'][t], ot['toff'][t + 1]) for t in range(mt.Ntrans)
                               if not p['trans'][t, 3]])
        assert np.array_equal(g, og[rows])


def test_synthetic_jitter_is_deterministic_and_mild():
    from lightspinner_b200 import synth
    p, r = load_golden('c1_falc_ca')
    a, b = synth.jitter_problem(p, 7), synth.jitter_problem(p, 7)
    assert np.array_equal(a['phi'], b['phi']) and np.array_equal(a['bg_chi'], b['bg_chi'])
    c = synth.jitter_problem(p, 8)
    assert not np.array_equal(a['bg_chi'], c['bg_chi'])
    assert np.all(np.abs(np.log(a['bg_chi'] / p['bg_chi'])) < 0.25)


@pytest.mark.reference
def test_context_host_mirror_against_live_reference():
    """With /root/reference present: the drop-in fed the REAL reference objects flattens to exactly what the
    reference's own Context holds (C, Nblue, damping parameters, ...)."""
    from oracle.refharness import reference_available, load_reference, build_falc_setup
    if not reference_available():
        pytest.skip('Lightspinner reference not present')
    from oracle.refharness.extract import problem_from_reference_context
    from lightspinner_b200.context import Context
    import copy
    ref = load_reference()
    atmos, spect, eqPops, bg = build_falc_setup(active=('Ca',), nrays=3)
    eq2 = copy.deepcopy(eqPops)
    rctx = ref['rh_method'].Context(atmos, spect, eqPops, bg)
    pr = problem_from_reference_context(rctx)
    ctx = Context(atmos, spect, eq2, bg, _host_only=True)
    for k in KEYS:
        assert np.array_equal(np.asarray(ctx._problem[k]), np.asarray(pr[k])), k


def test_stock_kernel_instances_cover_the_fixture_models():
    """tools/gen_spec_instances.py and mali_api.cu must agree on the structure keys: every tile of the fixture
    models has an ahead-of-time instance (the GPU twin checks model_info()['generic_tiles'] == 0)."""
    from lightspinner_b200 import specialize
    have = set(specialize.stock_keys())
    assert len(have) >= 40
    for name in ['c1_falc_ca', 'c2_falc_cah', 'c1v_jitter_ca3', 'rf_k40p']:
        p, _ = load_golden(name)
        assert set(specialize.tile_structures(p)) <= have, name
        assert specialize.library_for(p) is None


def test_batch_context_rejects_columns_of_different_models():
    """BatchContext checks, before touching the GPU, that its columns share one radiative model."""
    from lightspinner_b200.context import BatchContext
    p, _ = load_golden('c1_falc_ca')
    q = select_rays(p, [0, 2, 4])
    with pytest.raises(ValueError, match='share'):
        BatchContext([fake_reference_objects(p), fake_reference_objects(q)])
    with pytest.raises(ValueError):
        BatchContext([])


@pytest.mark.reference
def test_atom_and_eos_tables_from_live_reference_objects():
    """With /root/reference present: the model-level tables of the device-side set-up built from the REAL reference
    objects (AtomTables.from_models, EosTables.from_witt) equal the ones built from the committed fixtures
    (tests/golden/setup_inputs.npz, eos.npz) -- the path a user of the reference takes and the path the GPU tests take
    are the same data."""
    from oracle.refharness import reference_available, load_reference
    if not reference_available():
        pytest.skip('Lightspinner reference not present')
    import os
    from helpers import GOLDEN, load_setup_inputs
    from lightspinner_b200.atoms import AtomTables
    from lightspinner_b200.eos import EosTables
    ref = load_reference()
    models = [ref['rh_atoms'].H_6_atom(), ref['rh_atoms'].CaII_atom()]
    table = ref['atomic_set'].RadiativeSet(models).atomicTable
    live = AtomTables.from_models(models, table)
    atoms, _ = load_setup_inputs('c2_falc_cah')
    fix = AtomTables.from_arrays([dict(atoms['H']), dict(atoms['CA'])])
    for k in ('Nlevel', 'dE', 'gi0', 'dZ', 'nDebye', 'g', 'vTherm', 'coll', 'knots', 'coef', 'fill', 'par'):
        assert np.array_equal(getattr(live, k), getattr(fix, k)), k
    assert live.c1 == fix.c1 and live.c2 == fix.c2
    import witt as W
    z = np.load(os.path.join(GOLDEN, 'eos.npz'))
    e_live = EosTables.from_witt(W.witt(), ref['atomic_table'].get_global_atomic_table().weightPerH)
    e_fix = EosTables.from_arrays(z)
    for k in ('tpf', 'pf', 'eion', 'stage_off'):
        assert np.array_equal(getattr(e_live, k), getattr(e_fix, k)), k
    # witt.__init__ renormalises the class-level abundances in place on every instantiation: equal to rounding only
    assert np.allclose(e_live.abund, e_fix.abund, rtol=1e-14, atol=0)
    for k in ('avw', 'rho_from_H', 'ab_others', 'saha_fac', 'amu_wph', 'cm3', 'thomson'):
        assert abs(getattr(e_live, k) / getattr(e_fix, k) - 1.0) < 1e-14, k
