"""GPU parity on ragged / edge-case problems (mutated fixtures) against the oracle, and the generic-kernel path.

The oracle is pinned to the reference on the unmutated fixtures (tests/test_oracle_golden.py); here it is the
checker for shapes the reference fixtures do not cover: odd Nspace (no TMA staging), a single ray (32 wavelengths
per warp), 2 and 4 rays, no radiative transitions at all, a handful of transitions, and every tile forced through
the generic (non-specialised) kernel."""
import os

import numpy as np
import pytest

from helpers import drop_depth, load_golden, only_transitions, relerr, select_rays

pytestmark = pytest.mark.gpu


def gamma_err(G, Gref):
    scale = np.max(np.abs(Gref), axis=1, keepdims=True)
    scale[scale == 0] = 1.0
    return float(np.max(np.abs(G - Gref) / scale))


def run_both(p, oracle, niter=6, generic=False):
    from lightspinner_b200.engine import MaliEngine
    if generic:
        os.environ['MALI_NO_SPEC'] = '1'
    try:
        eng = MaliEngine(p, 1, arith='exact')     # the 1e-13 per-call bars below are the exact mode's
    finally:
        os.environ.pop('MALI_NO_SPEC', None)
    info = eng.model_info()
    eng.upload([p])
    oc = oracle.OracleContext(p)
    worst = dict(J=0.0, I=0.0, G=0.0, n=0.0)
    for it in range(1, niter + 1):
        eng.set_n(0, oc.n)
        dJ = float(eng.formal_sol_gamma_matrices()[0])
        dJo = oc.formal_sol_gamma_matrices()
        assert abs(dJ - dJo) <= 1e-11 * max(1.0, abs(dJo))
        worst['J'] = max(worst['J'], relerr(eng.J(0), oc.J))
        worst['I'] = max(worst['I'], relerr(eng.I(0), oc.I))
        worst['G'] = max(worst['G'], gamma_err(eng.Gamma(0), oc.Gamma))
        if it > 3:
            eng.stat_equil()
            oc.stat_equil(use_scipy=True)
            worst['n'] = max(worst['n'], relerr(eng.n(0), oc.n))
    eng.close()
    return worst, info


def check(w):
    assert w['J'] < 1e-13 and w['I'] < 1e-13 and w['G'] < 1e-11 and w['n'] < 1e-9, w


@pytest.mark.parametrize('name', ['c1_falc_ca', 'c2_falc_cah', 'c1v_jitter_ca3'])
def test_generic_kernel_path(oracle, name):
    """MALI_NO_SPEC: every tile goes through fs_gamma_kernel (runtime slot loops, shared-memory level sums)."""
    p, _ = load_golden(name)
    w, info = run_both(p, oracle, generic=True)
    assert info['spec_tiles'] == 0 and info['generic_tiles'] == info['ntile']
    check(w)


@pytest.mark.parametrize('name', ['c1_falc_ca', 'c2_falc_cah', 'c1v_jitter_ca3'])
def test_fixture_models_are_fully_specialised(oracle, name):
    p, _ = load_golden(name)
    from lightspinner_b200.engine import MaliEngine
    eng = MaliEngine(p, 1)
    info = eng.model_info()
    eng.close()
    assert info['generic_tiles'] == 0 and info['spec_tiles'] == info['ntile'], info
    assert info['tma'] == 1


def test_odd_nspace(oracle):
    """81 depth points: the last ring group of the specialised kernels holds a single depth step."""
    p, _ = load_golden('c1_falc_ca')
    q = drop_depth(p, 40)
    w, info = run_both(q, oracle)
    assert info['generic_tiles'] == 0
    check(w)


@pytest.mark.parametrize('rays', [[2], [0, 4], [0, 1, 3, 4]])
def test_other_ray_counts(oracle, rays):
    """1, 2 and 4 rays: 32 / 16 / 8 wavelengths per warp; no ahead-of-time instance -> generic kernel."""
    p, _ = load_golden('c1_falc_ca')
    q = select_rays(p, rays)
    w, info = run_both(q, oracle)
    check(w)


def test_no_transitions_at_all(oracle):
    """Pure background: no active transition on any wavelength (empty tiles everywhere)."""
    p, _ = load_golden('c1_falc_ca')
    q = only_transitions(p, [])
    w, info = run_both(q, oracle)
    assert w['J'] < 1e-13 and w['I'] < 1e-13 and w['G'] == 0.0, w


def test_few_transitions(oracle):
    """One line + one continuum of CaII: most tiles empty, ragged partial tiles at the ends of the line."""
    p, _ = load_golden('c1_falc_ca')
    q = only_transitions(p, [0, 5])
    w, info = run_both(q, oracle, niter=5)
    check(w)


def test_three_depth_points_minimum(oracle):
    """Nspace = 3 is the smallest atmosphere the reference's sweep accepts (one interior point)."""
    p, _ = load_golden('c1_falc_ca')
    q = p
    for k in range(int(p['Nspace']) - 1, 2, -1):
        q = drop_depth(q, 1)
    assert int(q['Nspace']) == 3
    w, info = run_both(q, oracle, niter=3)
    assert w['J'] < 1e-12 and w['I'] < 1e-12 and w['G'] < 1e-10, w


def test_limits_are_reported_not_crashed():
    from lightspinner_b200 import _capi
    from lightspinner_b200.engine import MaliEngine
    p, _ = load_golden('c1_falc_ca')
    q = drop_depth(drop_depth(p, 1), 1)
    for k in range(int(q['Nspace']) - 1, 1, -1):
        q = drop_depth(q, 1)
    assert int(q['Nspace']) == 2
    with pytest.raises(_capi.MaliError, match='3 depth points'):
        MaliEngine(q, 1)


def test_singular_system_raises_linalgerror():
    """A zeroed Gamma with zero populations is singular: the reference raises LinAlgError (rh_method.py:739)."""
    import torch
    from lightspinner_b200.engine import MaliEngine
    p, _ = load_golden('c1_falc_ca')
    eng = MaliEngine(p, 1)
    eng.upload([p])
    eng.formal_sol_gamma_matrices()
    eng.t_Gamma.zero_()
    with pytest.raises(np.linalg.LinAlgError):
        eng.stat_equil()
    eng.close()


def test_runtime_specialisation_builds_model_specific_kernels(oracle):
    """A model whose tile structures have no ahead-of-time instance (2 rays) gets a model-specific library built
    with nvcc (lightspinner_b200.specialize) and then runs entirely on specialised kernels, same results."""
    from lightspinner_b200 import specialize
    from lightspinner_b200.engine import MaliEngine
    p, _ = load_golden('c1_falc_ca')
    q = select_rays(p, [0, 4])
    assert set(specialize.tile_structures(q)) - set(specialize.stock_keys())
    eng = MaliEngine(q, 1, specialize=True, arith='exact')
    info = eng.model_info()
    assert info['generic_tiles'] == 0 and info['spec_tiles'] == info['ntile'], info
    eng.upload([q])
    oc = oracle.OracleContext(q)
    for it in range(1, 6):
        eng.set_n(0, oc.n)
        eng.formal_sol_gamma_matrices()
        oc.formal_sol_gamma_matrices()
        if it > 3:
            eng.stat_equil()
            oc.stat_equil(use_scipy=True)
    assert relerr(eng.J(0), oc.J) < 1e-13 and relerr(eng.I(0), oc.I) < 1e-13 and relerr(eng.n(0), oc.n) < 1e-9
    eng.close()


def test_synthetic_batch_properties_at_bench_scale(oracle):
    """The bench workload (synthetic CaII+H columns, lightspinner_b200/synth.py) at a quarter of the per-GPU batch:
    size-independent properties of the path plus a spot check of sampled columns against the oracle.
      * rate conservation: every column of Gamma sums to zero (rh_method.py:698-703);
      * a batch split in two launches gives bit-identical results to one launch;
      * sampled columns agree with the oracle run on the same synthetic inputs;
      * all populations stay positive and finite."""
    import torch
    from lightspinner_b200 import synth
    from lightspinner_b200.engine import MaliEngine
    from lightspinner_b200.tables import pack_column
    p, _ = load_golden('c2_falc_cah')
    ncol, iters = 256, 5
    dev = torch.device('cuda', 0)

    def run(splits):
        eng = MaliEngine(p, ncol, max_upload_chunk=ncol)
        hp = int(eng.lay.hostpack)
        base_pack = torch.from_numpy(pack_column(eng.mt, eng.lay, p)).to(dev)
        staging = torch.empty(ncol * hp, dtype=torch.float64, device=dev)
        synth.jitter_staging(staging, eng.lay, eng.mt, base_pack, list(range(ncol)))
        eng.repack_from_staging(staging, 0, ncol)
        for it in range(1, iters + 1):
            for c0, nc in splits:
                eng.formal_sol_gamma_async(c0, nc)
                if it > 3:
                    eng.stat_equil_async(c0, nc)
        torch.cuda.synchronize()
        out = (eng.t_pops.cpu().numpy().copy(), eng.t_J.cpu().numpy().copy(), eng.t_I.cpu().numpy().copy(),
               eng.t_Gamma.cpu().numpy().copy())
        samples = {c: (eng.J(c), eng.I(c), eng.n(c)) for c in (0, 101, 255)}
        lay, mt = eng.lay, eng.mt
        eng.close()
        return out, samples, lay, mt

    (n1, J1, I1, G1), samples, lay, mt = run([(0, ncol)])
    (n2, J2, I2, G2), _, _, _ = run([(0, 128), (128, 128)])
    assert np.array_equal(n1, n2) and np.array_equal(J1, J2) and np.array_equal(I1, I2) and np.array_equal(G1, G2)
    assert np.all(np.isfinite(n1)) and np.all(n1 > 0) and np.all(np.isfinite(I1))
    # conservation, per atom: sum over the first index of Gamma[i][j][k] vanishes to rounding
    N = mt.Nspace
    G = G1.reshape(ncol, -1)
    for a in range(mt.Natom):
        NL = int(mt.Nlevel[a])
        Ga = G[:, int(mt.g2off[a]) * N:int(mt.g2off[a + 1]) * N].reshape(ncol, NL, NL, N)
        assert np.max(np.abs(Ga.sum(axis=1))) <= 1e-9 * np.max(np.abs(Ga))
    # sampled columns against the oracle on the same synthetic inputs
    for c, (J, I, n) in samples.items():
        oc = oracle.OracleContext(synth.jitter_problem(p, c))
        for it in range(1, iters + 1):
            oc.formal_sol_gamma_matrices()
            if it > 3:
                oc.stat_equil(use_scipy=False)
        assert relerr(I, oc.I) < 1e-9 and relerr(J, oc.J) < 1e-9 and relerr(n, oc.n) < 1e-8, c
