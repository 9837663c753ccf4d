"""Device-side column set-up (lte_pops atomic_set.py:105-145, compute_collisions rh_method.py:474-487
over collisional_rates.py:36-96, v_broad atomic_model.py:241-245, continuum g_ij rh_method.py:453-454 ->
mali_model_set_atoms + mali_setup_columns) against what the unmodified reference computed for every column fixture, and
the MALI iteration run on a column set up that way.

Bars: the table interpolants are evaluated with scipy's own de Boor recurrence (bit-identical on the host); what
differs from the reference is the device's exp / pow (<= 2 ulp) -- nStar, C, vBroad and g_ij must agree to 1e-13."""
import numpy as np
import pytest
import torch

from helpers import load_golden, load_setup_inputs, relerr

pytestmark = pytest.mark.gpu

FIXTURES = ['c1_falc_ca', 'c2_falc_cah', 'c1v_jitter_ca3', 'c2v_jitter_cah_0', 'rf_k40p', 'stress_r10_d512']
TOL = 1e-13


def atom_tables(p):
    from lightspinner_b200.atoms import AtomTables
    atoms, _ = load_setup_inputs('c1_falc_ca')
    return AtomTables.from_arrays([dict(atoms[str(s).strip().upper()]) for s in p['atom_names']])


def with_atmosphere(p, name):
    _, col = load_setup_inputs(name)
    q = dict(p)
    q['ne'] = col['ne']
    return q


@pytest.mark.parametrize('name', FIXTURES)
def test_device_setup_matches_the_reference(name):
    from lightspinner_b200.engine import MaliEngine
    p, _ = load_golden(name)
    q = with_atmosphere(p, name)
    ref = MaliEngine(p, 1)
    ref.upload([p])                         # every table from the reference's own host arrays
    want = ref.t_colconst.cpu().numpy().copy()
    eng = MaliEngine(p, 1)
    eng.set_atoms(atom_tables(p))
    eng.upload_atmos([q])                   # LTE populations, C, vBroad, g_ij and the profiles formed on the device
    got = eng.t_colconst.cpu().numpy()
    N = int(p['Nspace'])
    nStar = eng.nStar.cpu().numpy().reshape(-1, N)
    e_ns = relerr(nStar, p['nStar'])
    e_vb = relerr(eng._last_vBroad.cpu().numpy()[0], p['vBroad'])
    nz = want != 0
    assert np.array_equal(got[~nz], want[~nz])          # nothing appears where the reference has nothing
    # line-profile entries carry the Voigt function's own 1e-12 bar (tests/test_gpu_device_phi.py); everything else
    # -- C, g_ij, heights, background -- the 1e-13 of this file.  Measure both over the whole block:
    e_all = float(np.max(np.abs(got[nz] - want[nz]) / np.abs(want[nz])))
    print('%s: nStar %.2e  vBroad %.2e  colconst %.2e' % (name, e_ns, e_vb, e_all))
    assert e_ns < TOL and e_vb < TOL
    assert e_all < 1e-12
    if np.array_equal(np.asarray(p['n']), np.asarray(p['nStar'])):      # a column that starts from LTE
        assert np.array_equal(eng.n(0), nStar)
    ref.close()
    eng.close()


def test_collisional_rates_block(oracle):
    """C alone, read back through Gamma: with no radiative transitions Gamma = C + its diagonal fix-up."""
    from helpers import only_transitions
    from lightspinner_b200.engine import MaliEngine
    p, _ = load_golden('c2_falc_cah')
    q = with_atmosphere(only_transitions(p, []), 'c2_falc_cah')
    eng = MaliEngine(q, 1)
    eng.set_atoms(atom_tables(p))
    eng.upload_atmos([q])
    eng.formal_sol_gamma_matrices()
    G = eng.Gamma(0)
    C = np.asarray(p['C'])
    N = int(p['Nspace'])
    o = 0
    for NL in (int(x) for x in p['Nlevel']):
        Ga, Ca = G[o:o + NL * NL].reshape(NL, NL, N), C[o:o + NL * NL].reshape(NL, NL, N)
        off = ~np.eye(NL, dtype=bool)
        assert relerr(Ga[off], Ca[off]) < TOL
        o += NL * NL
    eng.close()


def test_iteration_on_a_device_set_up_column():
    """CaII/FALC set up on the device from T, ne, nTotal, vturb: the reference's 46 iterations, I and n within 1e-10."""
    from lightspinner_b200.engine import MaliEngine
    p, r = load_golden('c1_falc_ca')
    eng = MaliEngine(p, 1)
    eng.set_atoms(atom_tables(p))
    eng.upload_atmos([with_atmosphere(p, 'c1_falc_ca')])
    eng.reset_iteration_state()
    for _ in range(8):
        eng.iterate_async(16)
        if bool((eng.t_done != 0).all().item()):
            break
    eng.raise_on_faults()
    assert int(eng.t_iter.cpu()[0]) == int(r['niter'])
    assert relerr(eng.I(0), r['final_I']) < 1e-10 and relerr(eng.n(0), r['final_n']) < 1e-10
    eng.close()


def test_warm_started_column_keeps_its_populations():
    """response_fn.py:33: start_from_lte=False leaves the caller's populations in place."""
    from lightspinner_b200.engine import MaliEngine
    p, r = load_golden('rf_k40p')
    eng = MaliEngine(p, 1)
    eng.set_atoms(atom_tables(p))
    eng.upload_atmos([with_atmosphere(p, 'rf_k40p')], start_from_lte=False)
    assert np.array_equal(eng.n(0), np.asarray(p['n']))
    eng.reset_iteration_state()
    eng.iterate_async(16)
    torch.cuda.synchronize()
    assert int(eng.t_iter.cpu()[0]) == int(r['niter'])
    assert relerr(eng.I(0), r['final_I']) < 1e-10
    eng.close()
