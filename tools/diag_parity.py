"""Diagnostic (GPU box): where does the population difference to the reference come from?
Separates (a) Gamma summation-order noise from (b) the linear solve."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from helpers import load_golden, relerr
from lightspinner_b200.engine import MaliEngine
from oracle import mali_oracle as mo

p, r = load_golden(sys.argv[1] if len(sys.argv) > 1 else 'c1_falc_ca')
eng = MaliEngine(p, 1); eng.upload([p])
oc = mo.OracleContext(p)
for it in range(1, 9):
    eng.set_n(0, oc.n)
    eng.formal_sol_gamma_matrices(); oc.formal_sol_gamma_matrices()
    G, Go = eng.Gamma(0), oc.Gamma
    nz = Go != 0
    rel = np.abs(G[nz] - Go[nz]) / np.abs(Go[nz])
    print('it %d  J err %.2e  I err %.2e  Gamma: max entrywise rel %.2e  median %.2e' %
          (it, relerr(eng.J(0), oc.J), relerr(eng.I(0), oc.I), rel.max(), np.median(rel)))
    if it > 3:
        nbefore = oc.n.copy()
        # (b) GPU solver on the oracle's exact Gamma
        eng.t_Gamma.copy_(torch.from_numpy(Go.reshape(-1)).to(eng.device))
        eng.set_n(0, nbefore)
        eng.stat_equil(); n_b = eng.n(0).copy()
        # (a) GPU solver on the GPU's own Gamma
        eng.t_Gamma.copy_(torch.from_numpy(G.reshape(-1)).to(eng.device))
        eng.set_n(0, nbefore)
        eng.stat_equil(); n_a = eng.n(0).copy()
        oc.stat_equil(use_scipy=True)
        print('      n err: own Gamma %.2e   oracle Gamma + GPU solver %.2e' % (relerr(n_a, oc.n), relerr(n_b, oc.n)))
