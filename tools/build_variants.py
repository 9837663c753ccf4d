#!/usr/bin/env python
"""Kernel-tuning helper: builds variants of libmali_b200.so with extra -D options (two at a time), named
libmali_b200_<tag>.so; tools/gpu_variants.sh then benches each one on the GPU (MALI_LIB_NAME selects the library).

    python tools/build_variants.py occ2_14:MALI_OCC2=14 nst2:MALI_NST=2 fma:MALI_FMA_GAMMA=1,MALI_OCC2=14
"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lightspinner_b200 import build as b  # noqa: E402


def one(spec):
    tag, _, defs = spec.partition(':')
    lib = os.path.join(b.LIBDIR, 'libmali_b200_%s.so' % tag)
    b.build(force=True, lib=lib, defines=[d for d in defs.split(',') if d])
    return lib


if __name__ == '__main__':
    with ThreadPoolExecutor(max_workers=2) as ex:
        for lib in ex.map(one, sys.argv[1:]):
            print('built', lib)
