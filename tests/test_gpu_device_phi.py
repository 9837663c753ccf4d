"""Device compute_phi (rh_method.py:198-243 -> mali_compute_phi) against the profiles the
unmodified reference produced (tests/golden: scipy.special.wofz on the host), and the MALI iteration run on them."""
import numpy as np
import pytest
import torch

from helpers import load_golden, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng_mod():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from lightspinner_b200 import engine
    return engine


@pytest.mark.parametrize('name', ['c1_falc_ca', 'c2_falc_cah', 'c1v_jitter_ca3', 'rf_k40p'])
def test_device_profiles_match_the_reference_tables(eng_mod, name):
    """Every entry of the device tables (Vij rows, wavelength-weight fields) built from device profiles vs the same
    tables built from the reference's phi / wphi: 1e-12 relative (wofz itself is good to ~1e-13)."""
    p, _ = load_golden(name)
    a = eng_mod.MaliEngine(p, 1)
    a.upload([p])
    ref = a.t_colconst.cpu().numpy().copy()
    b = eng_mod.MaliEngine(p, 1)
    b.upload_device_phi([p])
    got = b.t_colconst.cpu().numpy()
    size = int(a.mt.Nspace) * int(a.model_info()['row_stride'])      # the tile table is the tail of a column's block,
    off = int(a.lay.colconst) - size - (-size) % 16                   # which is padded to a multiple of 16 doubles
    r, g = ref[off:off + size], got[off:off + size]
    nz = r != 0
    assert np.array_equal(g[~nz], r[~nz])                 # inactive / padding entries stay zero
    assert float(np.max(np.abs(g[nz] - r[nz]) / np.abs(r[nz]))) < 1e-12     # measured: 2.5e-14
    a.close()
    b.close()


def test_iteration_on_device_profiles_matches_the_reference(eng_mod):
    """CaII/FALC to convergence on device-computed profiles: same iteration count, I and n within 1e-10."""
    p, r = load_golden('c1_falc_ca')
    e = eng_mod.MaliEngine(p, 1)
    e.upload_device_phi([p])
    e.reset_iteration_state()
    e.iterate_async(64)
    torch.cuda.synchronize()
    assert int(e.t_iter.cpu()[0]) == int(r['niter'])
    assert relerr(e.I(0), r['final_I']) < 1e-10
    assert relerr(e.n(0), r['final_n']) < 1e-10
    e.close()


def test_dropin_context_with_device_profiles(eng_mod):
    """Context forms its line profiles on the device; the test.py loop converges in the reference's 46 iterations
    to the same I, J, n; trans.phi / trans.wphi are still readable (read back from the device tables on first use)
    and agree with the reference's profiles to 1e-12."""
    from helpers import fake_reference_objects
    from lightspinner_b200 import Context
    p, r = load_golden('c1_falc_ca')
    atmos, spect, eqPops, bg = fake_reference_objects(p)
    ctx = Context(atmos, spect, eqPops, bg)
    assert all(t._profile is None for t in ctx.activeAtoms[0].trans)
    dJ, dPops, i = 1.0, 1.0, 0
    while dJ > 2e-3 or dPops > 1e-3:
        i += 1
        dJ = ctx.formal_sol_gamma_matrices()
        if i > 3:
            dPops = ctx.stat_equil()
    assert i == int(r['niter'])
    assert relerr(ctx.I, r['final_I']) < 1e-10 and relerr(ctx.J, r['final_J']) < 1e-10
    assert relerr(eqPops['CA'].n, r['final_n']) < 1e-10
    t = ctx.activeAtoms[0].trans[0]
    Nlam = int(p['trans'][0, 5])
    assert t.phi.shape == (Nlam, 5, 2, 82) and t.wphi.shape == (82,)
    ref_phi = np.asarray(p['phi'])[int(p['phioff'][0]):int(p['phioff'][0]) + Nlam * 5 * 2 * 82].reshape(Nlam, 5, 2, 82)
    assert relerr(t.phi, ref_phi) < 1e-12 and relerr(t.wphi, p['wphi'][0]) < 1e-12
    assert np.array_equal(t.wlambda(), ctx._engine.mt.wlambda[:Nlam]) and t.wlambda(3) == ctx._engine.mt.wlambda[3]
    ctx.close()


def test_batch_context_runs_the_response_function_columns(eng_mod):
    """BatchContext: the base column and two response-function columns (different atmospheres, warm-started through
    the eqPops alias) solved as one batch; every column converges in the reference's iteration count to the
    reference's I and n, and the per-column Context views and eqPops aliases see the results."""
    from helpers import fake_reference_objects
    from lightspinner_b200 import BatchContext
    names = ['c1_falc_ca', 'rf_k40p', 'rf_k10m']
    gold = [load_golden(nm) for nm in names]
    cols = [fake_reference_objects(g[0]) for g in gold]
    batch = BatchContext(cols)
    its = batch.iterate()
    assert len(batch) == 3
    for k, (p, r) in enumerate(gold):
        assert int(its[k]) == int(r['niter']), (names[k], its[k])
        assert relerr(batch[k].I, r['final_I']) < 1e-10
        assert relerr(cols[k][2]['CA'].n, r['final_n']) < 1e-10          # the eqPops alias of column k
        assert batch[k].J.shape == r['final_J'].shape
    # the per-call forms, batched and per column
    dJ = batch.formal_sol_gamma_matrices()
    assert dJ.shape == (3,) and np.all(dJ < 2e-3)
    dP = batch.stat_equil()
    assert dP.shape == (3,) and np.all(dP < 1e-3)
    d1 = batch[1].formal_sol_gamma_matrices()
    assert isinstance(d1, float) and d1 < 2e-3
    batch.close()
