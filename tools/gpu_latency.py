#!/usr/bin/env python
"""Latency-bound cases on one GPU (BASELINE configs 1-3): one CaII/FALC column to convergence through the
device-resident loop (graph replay vs plain launches: MALI_NO_GRAPH=1), and the 164-column response-function batch.

    python tools/gpu_latency.py            # prints one JSON line
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from helpers import load_golden, relerr  # noqa: E402
from lightspinner_b200.engine import MaliEngine  # noqa: E402


def converge(eng, ncol, chunk=16, cap=128):
    eng.reset_iteration_state()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    done = 0
    while done < cap:
        eng.iterate_async(chunk, ncol=ncol)
        done += chunk
        if bool((eng.t_done[:ncol] != 0).all().item()):
            break
    torch.cuda.synchronize()
    return time.perf_counter() - t0


def main():
    out = {'graph': os.environ.get('MALI_NO_GRAPH') is None}
    for name in ('c1_falc_ca', 'c2_falc_cah'):
        p, r = load_golden(name)
        eng = MaliEngine(p, 1)
        ts = []
        for rep in range(5):
            eng.upload([p])
            ts.append(converge(eng, 1))
        eng.raise_on_faults()
        out[name] = {'iterations': int(eng.t_iter.cpu()[0]), 'reference_iterations': int(r['niter']),
                     'seconds_best': min(ts), 'seconds_all': ts, 'rel_err_n': relerr(eng.n(0), r['final_n']),
                     'rel_err_I': relerr(eng.I(0), r['final_I'])}
        eng.close()
    gold = [load_golden(nm) for nm in ('rf_k40p', 'rf_k10m')]
    n_rf = 164
    probs = [gold[c % 2][0] for c in range(n_rf)]
    eng = MaliEngine(probs[0], n_rf, max_upload_chunk=n_rf)
    ts = []
    for rep in range(4):
        eng.upload_device_phi(probs)
        ts.append(converge(eng, n_rf, chunk=8))
    its = eng.t_iter.cpu().numpy()
    out['response_function'] = {'columns': n_rf, 'seconds_best_solve_only': min(ts), 'seconds_all': ts,
                                'iterations': sorted(set(int(x) for x in its)),
                                'matches': bool(all(int(its[c]) == int(gold[c % 2][1]['niter']) for c in range(n_rf)))}
    eng.close()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
