import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import load_golden
import lightspinner_b200.engine as E
p, r = load_golden('c1v_jitter_ca3')
N, S = 82, 287
runs = []
for rep in range(4):
    eng = E.MaliEngine(p, 1); eng.upload([p])
    eng.formal_sol_gamma_matrices()
    sc = eng.t_scratch.cpu().numpy().copy(); L = eng.lay
    runs.append(sc); eng.close()
half = L.scratch // 2
ref = runs[0]
for i in range(1, 4):
    d = np.argwhere(runs[i] != ref).ravel()
    print('run', i, 'differs in', len(d), 'entries')
    if len(d):
        dn = d[d < half]; up = d[d >= half] - half
        for name, idx in (('down', dn), ('up', up)):
            j = idx[idx < N * S]; g = idx[idx >= N * S]
            print(' ', name, 'Jpart entries', len(j), 'k range', (j // S).min() if len(j) else None, (j // S).max() if len(j) else None,
                  'la', sorted(set((j % S).tolist()))[:12], 'Gamma-part entries', len(g))
            if len(j):
                q = j[0]; print('   example q', q, 'k', q // S, 'la', q % S, ref[q + (half if name == 'up' else 0)], runs[i][q + (half if name == 'up' else 0)])
