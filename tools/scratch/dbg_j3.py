import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import load_golden
import lightspinner_b200.engine as E
p, r = load_golden('c1v_jitter_ca3')
for rep in range(3):
    eng = E.MaliEngine(p, 1); eng.upload([p])
    cc = eng.t_colconst.cpu().numpy().copy()
    eng.formal_sol_gamma_matrices()
    cc2 = eng.t_colconst.cpu().numpy().copy()
    d = np.argwhere(cc2 != cc).ravel()
    print(rep, 'changed', len(d), 'nonzero before FS among them', np.count_nonzero(cc[d]), 'max', np.abs(cc[d]).max())
    eng.upload([p])
    cc3 = eng.t_colconst.cpu().numpy().copy()
    print('   after re-upload: nonzero at J-dagger positions', np.count_nonzero(cc3[d]), ' identical to first upload:', np.array_equal(cc3, cc))
    eng.close()
