#!/usr/bin/env python
"""Times the device-side set-up stages for a batch of config-4 columns given by their thermodynamic state
(mali_upload_columns_thermo / mali_background / mali_setup_columns / mali_compute_phi), CUDA events per stage."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from helpers import load_golden, load_setup_inputs  # noqa: E402
from lightspinner_b200 import synth  # noqa: E402
from lightspinner_b200.atoms import AtomTables  # noqa: E402
from lightspinner_b200.engine import MaliEngine  # noqa: E402
from lightspinner_b200.eos import EosTables  # noqa: E402
from lightspinner_b200.tables import pack_column  # noqa: E402


def main():
    ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    p, _ = load_golden('c2_falc_cah')
    z = np.load(os.path.join(ROOT, 'tests', 'golden', 'eos.npz'))
    atoms, _ = load_setup_inputs('c2_falc_cah')
    eng = MaliEngine(p, ncol, max_upload_chunk=ncol)
    eng.set_eos(EosTables.from_arrays(z))
    eng.set_atoms(AtomTables.from_arrays([dict(atoms[str(s).strip().upper()]) for s in p['atom_names']]))
    N = int(p['Nspace'])
    lay, mt = eng.lay, eng.mt
    hpb = int(lay.hp_bg_chi)
    t0 = time.perf_counter()
    host = np.zeros((ncol, hpb))
    aux = {k: np.zeros((ncol, N)) for k in ('temperature', 'ne', 'nHTot', 'vturb', 'vlos', 'cmass')}
    ab = [atoms[str(s).strip().upper()]['abundance'] for s in p['atom_names']]
    for c in range(ncol):
        T, ne, vlos = synth.jitter_atmosphere(c, z['falc_T'], z['falc_ne'])
        q = dict(temperature=T, nTotal=np.stack([a * z['falc_nHTot'] for a in ab]))
        pack_column(mt, lay, q, out=host[c], with_background=False)
        aux['temperature'][c], aux['ne'][c], aux['vlos'][c] = T, ne, vlos
        aux['nHTot'][c], aux['vturb'][c], aux['cmass'][c] = z['falc_nHTot'], p['vturb'], z['falc_cmass']
    t_host = time.perf_counter() - t0
    dev = {k: torch.from_numpy(v).cuda() for k, v in aux.items()}
    aD = torch.from_numpy(np.broadcast_to(np.asarray(p['aDamp'])[None], (ncol,) + np.asarray(p['aDamp']).shape).copy()).cuda()
    pinned = torch.from_numpy(host.reshape(-1)).pin_memory()
    staging = torch.empty(ncol * lay.hostpack, dtype=torch.float64, device='cuda')
    work = torch.empty(ncol * N * 21, dtype=torch.float64, device='cuda')
    nStar = torch.empty(ncol * lay.sumNlevel * N, dtype=torch.float64, device='cuda')
    vB = torch.empty((ncol, mt.Natom, N), dtype=torch.float64, device='cuda')
    P = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    out = {'ncol': ncol, 'host_pack_seconds': t_host, 'h2d_bytes': int(pinned.numel() * 8 + sum(v.nbytes for v in aux.values()) + aD.numel() * 8)}
    for rep in range(2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        eng._check(eng.lib.mali_upload_columns_thermo(eng._handle, C.byref(eng.bufs), 0, ncol, P(pinned), P(staging), st))
        ev[1].record()
        eng._check(eng.lib.mali_background(eng._handle, C.byref(eng.bufs), 0, ncol, P(dev['temperature']), P(dev['ne']),
                                           P(dev['nHTot']), P(dev['cmass']), P(work), st))
        ev[2].record()
        eng._check(eng.lib.mali_setup_columns(eng._handle, C.byref(eng.bufs), 0, ncol, P(dev['temperature']), P(dev['ne']),
                                              P(dev['vturb']), P(nStar), P(vB), 1, st))
        ev[3].record()
        eng._check(eng.lib.mali_compute_phi(eng._handle, C.byref(eng.bufs), 0, ncol, P(aD), P(vB), P(dev['vlos']), st))
        ev[4].record()
        torch.cuda.synchronize()
        out['ms_upload_pack'], out['ms_background'], out['ms_setup'], out['ms_phi'] = (ev[i].elapsed_time(ev[i + 1]) for i in range(4))
    eng.reset_iteration_state()
    for _ in range(12):
        eng.iterate_async(16)
        if bool((eng.t_done != 0).all().item()):
            break
    its = eng.t_iter.cpu().numpy()
    out['iterations_min_max_mean'] = [int(its.min()), int(its.max()), float(its.mean())]
    out['iterations_col0_col1'] = [int(its[0]), int(its[1])]
    eng.raise_on_faults()
    print(json.dumps(out))


if __name__ == '__main__':
    main()
