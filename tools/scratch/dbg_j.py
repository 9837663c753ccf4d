import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import load_golden
import lightspinner_b200.engine as E
from oracle import mali_oracle as O
p, r = load_golden('c1v_jitter_ca3')
eng = E.MaliEngine(p, 1); eng.upload([p])
info = eng.model_info(); print(info)
cc = eng.t_colconst.cpu().numpy()
L = eng.lay
print('colconst', L.colconst, 'nan count', np.isnan(cc).sum())
oc = O.OracleContext(p)
eng.formal_sol_gamma_matrices(); oc.formal_sol_gamma_matrices()
J = eng.J(0); Jo = oc.J
rel = np.abs(J - Jo) / np.abs(Jo)
bad = np.argwhere(rel > 1e-14)
print('n bad', len(bad), 'la set', sorted(set(int(x) for x in bad[:, 0])))
for la in (40, 41, 42, 43, 110, 111):
    print(la, J[la, :4], Jo[la, :4])
I = eng.I(0); print('I err', np.abs(I - oc.I).max() / np.abs(oc.I).max())
cc2 = eng.t_colconst.cpu().numpy()
d = np.argwhere(cc2 != cc).ravel()
print('colconst entries changed by FS:', len(d), 'expected', J.size)
sc = eng.t_scratch.cpu().numpy()
N, S = 82, 287
half = L.scratch // 2
dn = sc[:N * S].reshape(N, S).T; up = sc[half:half + N * S].reshape(N, S).T
badla = sorted(set(int(x) for x in bad[:, 0]))
for la in badla[:3] + [0, 1]:
    print('la', la, 'down', dn[la, :3], 'up', up[la, :3], 'sum', (dn + up)[la, :3], 'oracle', Jo[la, :3])
