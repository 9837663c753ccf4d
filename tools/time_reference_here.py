#!/usr/bin/env python
"""Time the UNMODIFIED reference's own numpy/numba path in THIS container (where /root/reference is mounted) and
record it under profiles/ -- the GPU box has no reference, so bench.py's CPU arm there is the oracle's C port.

    python tools/time_reference_here.py            # writes profiles/r01_reference_numpy_timing.json
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refharness as rh  # noqa: E402  (test infrastructure; never imported by the product)


def main():
    out = {'where': 'build container (CPU only), single process, single thread',
           'reference': rh.REFERENCE_DIR, 'cases': []}
    for active, name in ((('Ca',), 'C1 CaII/FALC'), (('Ca', 'H'), 'C2 CaII+H/FALC')):
        t0 = time.perf_counter()
        atmos, spect, eqPops, bg = rh.build_falc_setup(active=active, nrays=5)
        ctx = rh.load_reference()['rh_method'].Context(atmos, spect, eqPops, bg)
        t_setup = time.perf_counter() - t0
        Nspect, Nrays, Nspace = ctx.I.shape[0], ctx.I.shape[1], ctx.J.shape[1]
        units = Nspect * Nrays * Nspace
        ctx.formal_sol_gamma_matrices()          # numba compile + first touch
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            ctx.formal_sol_gamma_matrices()
            ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        ctx.stat_equil()
        t_se = time.perf_counter() - t0
        fs = sorted(ts)[1]
        out['cases'].append({'case': name, 'Nspect': Nspect, 'Nrays': Nrays, 'Nspace': Nspace,
                             'units_per_iteration': units, 'setup_seconds': t_setup,
                             'formal_sol_gamma_matrices_seconds_median_of_3': fs, 'stat_equil_seconds': t_se,
                             'updates_per_s_one_core': units / (fs + t_se)})
        print(out['cases'][-1])
    # a multiprocessing.Pool over columns -- every worker builds its own reference Context (CaII+H /
    # FALC) and times formal_sol_gamma_matrices + stat_equil calls; aggregate updates/s over all cores of this container
    import multiprocessing as mp
    ncore = os.cpu_count() or 1
    with mp.get_context('fork').Pool(ncore) as pool:
        res = pool.map(_pool_worker, range(ncore))
    units = res[0][0]
    agg = sum(u * n / t for u, n, t in res)
    out['pool'] = {'processes': ncore, 'calls_per_process': res[0][1], 'updates_per_s_aggregate': agg,
                   'updates_per_s_per_core': agg / ncore, 'case': 'C2 CaII+H/FALC, one Context per process'}
    print(out['pool'])
    json.dump(out, open(os.path.join(ROOT, 'profiles', 'r02_reference_numpy_timing.json'), 'w'), indent=1)


def _pool_worker(i):
    atmos, spect, eqPops, bg = rh.build_falc_setup(active=('Ca', 'H'), nrays=5)
    ctx = rh.load_reference()['rh_method'].Context(atmos, spect, eqPops, bg)
    units = ctx.I.shape[0] * ctx.I.shape[1] * ctx.J.shape[1]
    for _ in range(4):                       # numba compile + the J-only iterations
        ctx.formal_sol_gamma_matrices()
    ncall = 3
    t0 = time.perf_counter()
    for _ in range(ncall):
        ctx.formal_sol_gamma_matrices()
        ctx.stat_equil()
    return units, ncall, time.perf_counter() - t0


if __name__ == '__main__':
    main()
