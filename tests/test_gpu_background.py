"""Device-side background / EOS / scale conversion (Background.compute_background_eos
background.py:21-53 over witt.py, AtmosphereConstructor.convert_scales atmosphere.py:70-112 -> mali_model_set_eos +
mali_background) against the numbers the unmodified reference produced (tests/golden/eos.npz), and the MALI iteration
of a column set up from its thermodynamic state alone.

The source the kernels compile (csrc/mali_eos.h) is bit-identical to the reference on the host (tests/test_eos_host.py);
on the device exp / log / pow differ by <= 2 ulp, and the EOS iterations stop at a relative 1e-5, so that rounding
enters the pressures at the 1e-15 level only: the bar is 1e-12 on chi, eta, sca, the pressures and the heights."""
import os

import numpy as np
import pytest
import torch

from helpers import load_golden, load_setup_inputs, relerr

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TOL = 1e-12


def engine_for(p, ncol=1):
    from lightspinner_b200.atoms import AtomTables
    from lightspinner_b200.engine import MaliEngine
    from lightspinner_b200.eos import EosTables
    z = np.load(os.path.join(HERE, 'golden', 'eos.npz'))
    atoms, _ = load_setup_inputs('c1_falc_ca')
    eng = MaliEngine(p, ncol)
    eng.set_eos(EosTables.from_arrays(z))
    eng.set_atoms(AtomTables.from_arrays([dict(atoms[str(s).strip().upper()]) for s in p['atom_names']]))
    return eng, z


def thermo_problem(p, z, case, abundances):
    q = {k: p[k] for k in ('Nspace', 'Nrays', 'Nspect', 'wavelength', 'muz', 'wmu', 'Nlevel', 'trans', 'linepar', 'alpha',
                           'vturb', 'vlos', 'aDamp', 'atom_names')}
    q['temperature'], q['ne'], q['nHTot'], q['cmass'] = (z['%s_%s' % (case, k)] for k in ('T', 'ne', 'nHTot', 'cmass'))
    q['nTotal'] = np.stack([a * q['nHTot'] for a in abundances])
    return q


@pytest.mark.parametrize('case,fixture', [('falc', 'c2_falc_cah'), ('jitter0', 'c2v_jitter_cah_0')])
def test_background_and_heights_match_the_reference(case, fixture):
    p, _ = load_golden(fixture)
    eng, z = engine_for(p)
    atoms, _ = load_setup_inputs(fixture)
    q = thermo_problem(p, z, case, [atoms[str(s).strip().upper()]['abundance'] for s in p['atom_names']])
    eng.upload_thermo([q])
    N, S = int(p['Nspace']), int(p['Nspect'])
    work = eng._last_work.cpu().numpy()[:N * 20].reshape(N, 20)
    e_pg, e_pe = relerr(work[:, 0], z[case + '_pgas']), relerr(work[:, 1], z[case + '_pe'])
    e_cc, e_parts = relerr(work[:, 2], z[case + '_chi_c']), relerr(work[:, 3:], z[case + '_partials'])
    # the tables the upload produced vs the same tables built from the reference's own background arrays and heights
    ref_p = dict(p)
    ref_p['bg_chi'], ref_p['bg_eta'] = z[case + '_chi'], z[case + '_eta']
    ref_p['bg_sca'] = np.broadcast_to(z[case + '_thomson'], (S, N)).copy()
    ref_p['height'] = z[case + '_height']
    from lightspinner_b200.engine import MaliEngine
    ref = MaliEngine(p, 1)
    ref.upload([ref_p])
    want, got = ref.t_colconst.cpu().numpy(), eng.t_colconst.cpu().numpy()
    nz = want != 0
    e_tab = float(np.max(np.abs(got[nz] - want[nz]) / np.abs(want[nz])))
    # heights: relative to the extent of the atmosphere (they pass through zero at tau_500 = 1)
    zz = got[:N]
    e_h = float(np.max(np.abs(zz - z[case + '_height'])) / np.max(np.abs(z[case + '_height'])))
    print('%s: pgas %.2e pe %.2e chi_c %.2e partials %.2e  tables %.2e heights %.2e' % (case, e_pg, e_pe, e_cc, e_parts, e_tab, e_h))
    assert e_pg < TOL and e_pe < TOL and e_cc < TOL and e_parts < TOL
    assert e_h < TOL
    assert not got[~nz].any()            # nothing appears where the reference has nothing
    assert e_tab < TOL                   # every table entry: background fields, C, g_ij, profiles (measured 8e-14)
    ref.close()
    eng.close()


def test_iteration_from_the_thermodynamic_state_alone():
    """CaII/FALC from (cmass, T, ne, nHTot, vturb): EOS, background, heights, LTE populations, collisional rates,
    profiles all on the device; the reference's 46 iterations, I and n within 1e-10 of its converged output."""
    p, r = load_golden('c1_falc_ca')
    eng, z = engine_for(p)
    atoms, _ = load_setup_inputs('c1_falc_ca')
    q = thermo_problem(p, z, 'falc', [atoms['CA']['abundance']])
    eng.upload_thermo([q])
    eng.reset_iteration_state()
    for _ in range(8):
        eng.iterate_async(16)
        if bool((eng.t_done != 0).all().item()):
            break
    eng.raise_on_faults()
    assert int(eng.t_iter.cpu()[0]) == int(r['niter'])
    e_I, e_n = relerr(eng.I(0), r['final_I']), relerr(eng.n(0), r['final_n'])
    print('from the thermodynamic state: rel err I %.2e  n %.2e' % (e_I, e_n))
    assert e_I < 1e-10 and e_n < 1e-10
    eng.close()
