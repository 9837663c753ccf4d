import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import load_golden
import lightspinner_b200.engine as E
from scipy.special import wofz
for name in ('c2_falc_cah', 'c1v_jitter_ca3', 'rf_k40p'):
    p, _ = load_golden(name)
    a = E.MaliEngine(p, 1); a.upload([p]); ref = a.t_colconst.cpu().numpy().copy()
    b = E.MaliEngine(p, 1); b.upload_device_phi([p]); got = b.t_colconst.cpu().numpy()
    off = int(a.lay.colconst) - int(a.mt.Nspace) * int(a.model_info()['row_stride'])
    r, g = ref[off:], got[off:]
    nz = r != 0
    rel = np.zeros_like(r); rel[nz] = np.abs(g[nz] - r[nz]) / np.abs(r[nz])
    print(name, 'max rel', rel.max(), 'zeros equal', np.array_equal(g[~nz], r[~nz]), 'prefix equal', np.array_equal(got[:off], ref[:off]))
    i = rel.argmax(); rs = int(a.model_info()['row_stride'])
    print('   worst at k', i // rs, 'row offset', i % rs, 'ref', r[i], 'got', g[i], ' n>1e-12:', int((rel > 1e-12).sum()), ' n>1e-11:', int((rel > 1e-11).sum()))
    print('   aDamp range', p['aDamp'][p['aDamp'] > 0].min(), p['aDamp'].max())
    a.close(); b.close()
