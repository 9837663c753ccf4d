"""Wavelength-sharded single column (BASELINE config 5 experiment): the sub-problems of lambda_shard.py against the
oracle on CPU -- single process and a world_size-2 gloo run of the Gamma all-reduce -- and against the CUDA path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import load_golden
from lightspinner_b200.lambda_shard import lambda_ranges, lambda_shard_problem

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def eng_mod():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from lightspinner_b200 import engine
    return engine


def test_lambda_ranges_cover_spectrum_on_tile_boundaries():
    for S in (1, 5, 287, 777, 2861):
        for world in (1, 2, 3, 8):
            for align in (1, 3, 6, 32):
                r = lambda_ranges(S, world, align)
                assert r[0][0] == 0 and r[-1][1] == S
                for (a0, a1), (b0, b1) in zip(r[:-1], r[1:]):
                    assert a1 == b0 and a0 <= a1 and a1 % align == 0


def test_work_balanced_ranges():
    """Ranges balanced by the per-wavelength work estimate: same covering / alignment rules, every rank gets at least one
    tile, and on the config-5 fixture the heaviest rank carries less than with equal wavelength counts."""
    from lightspinner_b200.lambda_shard import wavelength_cost
    rng = np.random.default_rng(3)
    for S in (7, 287, 2861):
        cost = 1.1 + rng.integers(0, 9, S)
        for world in (1, 2, 3, 8):
            for align in (1, 3, 6):
                if (S + align - 1) // align < world:
                    continue
                r = lambda_ranges(S, world, align, cost=cost)
                assert len(r) == world and r[0][0] == 0 and r[-1][1] == S
                for (a0, a1), (b0, b1) in zip(r[:-1], r[1:]):
                    assert a1 == b0 and a0 < a1 and a1 % align == 0
                assert r[-1][0] < r[-1][1]
    p, _ = load_golden('stress_r10_d512')
    cost = wavelength_cost(p)
    S = int(p['Nspect'])
    for world in (2, 4, 8):
        heavy = max(cost[a:b].sum() for a, b in lambda_ranges(S, world, 3, cost=cost))
        even = max(cost[a:b].sum() for a, b in lambda_ranges(S, world, 3))
        assert heavy < even and heavy < 1.2 * cost.sum() / world


@pytest.mark.parametrize('name,world', [('c1_falc_ca', 3), ('c2_falc_cah', 2)])
def test_shards_reproduce_the_full_formal_solution(oracle, name, world):
    p, _ = load_golden(name)
    full = oracle.OracleContext(p)
    full.formal_sol_gamma_matrices()
    G = None
    Lw = 32 // int(p['Nrays'])
    for r, (lo, hi) in enumerate(lambda_ranges(int(p['Nspect']), world, Lw)):
        q = lambda_shard_problem(p, lo, hi, keep_C=(r == 0))
        oc = oracle.OracleContext(q)
        oc.formal_sol_gamma_matrices()
        assert np.array_equal(oc.J, full.J[lo:hi])      # per-wavelength quantities: bit-identical
        assert np.array_equal(oc.I, full.I[lo:hi])
        G = oc.Gamma.copy() if G is None else G + oc.Gamma
    assert np.max(np.abs(G - full.Gamma)) <= 1e-13 * np.max(np.abs(full.Gamma))   # summation order only


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    from oracle import mali_oracle as O
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    p, _ = load_golden('c1_falc_ca')
    lo, hi = lambda_ranges(int(p['Nspect']), world, 6)[rank]
    oc = O.OracleContext(lambda_shard_problem(p, lo, hi, keep_C=(rank == 0)))
    hist = []
    from lightspinner_b200.lambda_shard import GammaExchange
    G = torch.zeros(oc.Gamma.size, dtype=torch.float64)
    dJ = torch.zeros(1, dtype=torch.float64)
    exchange = GammaExchange(G, dJ)      # the exchange of lambda_shard.LambdaShardedColumn itself (gloo here)
    for it in range(1, 7):           # the loop of test.py
        dJ[0] = oc.formal_sol_gamma_matrices()
        G.copy_(torch.from_numpy(np.ascontiguousarray(oc.Gamma)).view(-1))
        exchange()
        oc.Gamma[...] = G.numpy().reshape(oc.Gamma.shape)
        dP = oc.stat_equil(use_scipy=False) if it > 3 else None
        hist.append((float(dJ), dP))
    q.put((rank, oc.n.copy(), hist))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_iteration_matches_unsharded(oracle):
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = {r: (n, h) for r, n, h in (q.get(timeout=300) for _ in range(2))}
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p, _ = load_golden('c1_falc_ca')
    full = oracle.OracleContext(p)
    hist = []
    for it in range(1, 7):
        dJ = full.formal_sol_gamma_matrices()
        dP = full.stat_equil(use_scipy=False) if it > 3 else None
        hist.append((dJ, dP))
    assert np.array_equal(res[0][0], res[1][0])                       # replicas stay identical
    assert np.max(np.abs(res[0][0] - full.n) / np.abs(full.n)) < 1e-10
    for (a, b), (c, d) in zip(res[0][1], hist):
        assert abs(a - c) <= 1e-10 * max(1.0, abs(c))
        assert (b is None) == (d is None) and (b is None or abs(b - d) <= 1e-9 * max(1.0, abs(d)))


@pytest.mark.gpu
def test_gpu_shards_sum_to_the_unsharded_gamma(eng_mod):
    """Two wavelength shards run one after the other on one GPU: J and I bit-identical to the unsharded engine's
    slices, Gamma sums to the unsharded Gamma up to summation order."""
    p, _ = load_golden('c2_falc_cah')
    full = eng_mod.MaliEngine(p, 1)
    full.upload([p])
    full.formal_sol_gamma_matrices()
    Jf, If, Gf = full.J(0), full.I(0), full.Gamma(0)
    G = None
    for r, (lo, hi) in enumerate(lambda_ranges(int(p['Nspect']), 2, 32 // int(p['Nrays']))):
        q = lambda_shard_problem(p, lo, hi, keep_C=(r == 0))
        e = eng_mod.MaliEngine(q, 1)
        e.upload([q])
        e.formal_sol_gamma_matrices()
        assert np.array_equal(e.J(0), Jf[lo:hi])
        assert np.array_equal(e.I(0), If[lo:hi])
        G = e.Gamma(0).copy() if G is None else G + e.Gamma(0)
        e.close()
    assert np.max(np.abs(G - Gf)) <= 1e-12 * np.max(np.abs(Gf))
    full.close()
