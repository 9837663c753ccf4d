#!/usr/bin/env python
"""Target of the compute-sanitizer runs (profiles/r02_sanitizer_*.txt): a few MALI iterations of one fixture column
through the C ABI -- upload, formal solution + Gamma, statistical equilibrium, the device-resident loop.

    compute-sanitizer --tool racecheck python tools/sanitize_run.py c1_falc_ca
    MALI_NO_SPEC=1 compute-sanitizer --tool memcheck python tools/sanitize_run.py c2_falc_cah   # generic kernel
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from helpers import load_golden, relerr  # noqa: E402
from lightspinner_b200.engine import MaliEngine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'c1_falc_ca'
    p, r = load_golden(name)
    eng = MaliEngine(p, 2)
    eng.upload([p, p])
    for it in range(1, 6):
        eng.formal_sol_gamma_matrices()
        if it > 3:
            eng.stat_equil()
    e = relerr(eng.I(0), r['it5_I']) if 'it5_I' in r else float('nan')
    eng.upload_device_phi([p, p])
    eng.reset_iteration_state()
    eng.iterate_async(3)
    torch.cuda.synchronize()
    print('sanitize_run %s: info %s  rel err I after 5 iterations %.2e  finite %s' % (
        name, eng.model_info(), e, bool(np.isfinite(eng.n(1)).all())))
    eng.close()


if __name__ == '__main__':
    main()
