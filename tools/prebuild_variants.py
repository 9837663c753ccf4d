#!/usr/bin/env python
"""Builds, in the CPU container, the model-specific library variants the GPU tests ask for, so that the GPU box finds
them cached in lightspinner_b200/_lib/ (built .so files travel with the gpurun snapshot) instead of spending minutes
of GPU-box time in nvcc: the 2-ray CaII model of tests/test_gpu_edge_cases.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from helpers import load_golden, select_rays  # noqa: E402
from lightspinner_b200 import specialize  # noqa: E402

if __name__ == '__main__':
    for f in os.listdir(os.path.join(ROOT, 'lightspinner_b200', '_lib')):
        if f.startswith('libmali_b200_spec_') or f.startswith('spec_instances_'):
            os.remove(os.path.join(ROOT, 'lightspinner_b200', '_lib', f))      # stale variants of older kernels
    p, _ = load_golden('c1_falc_ca')
    print(specialize.library_for(select_rays(p, [0, 4]), verbose=True))
