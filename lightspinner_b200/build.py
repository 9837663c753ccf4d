"""In-tree build of libmali_b200.so with nvcc for sm_100a (no torch, no JIT cache).

    python -m lightspinner_b200.build [--force] [--verbose]

--fmad=false is part of the numerical contract (see csrc/mali_kernels.cuh); -lineinfo keeps the ncu source
page usable.  The built library lives in lightspinner_b200/_lib/ (git-ignored, travels with gpurun).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, '_lib')
LIB = os.path.join(LIBDIR, os.environ.get('MALI_LIB_NAME', 'libmali_b200.so'))
SOURCES = ['mali_api.cu']
DEPS = ['mali_api.cu', 'mali_kernels.cuh', 'mali_types.cuh', 'exp_table.inc', 'mali_solve.h', 'mali_voigt.h', 'mali_fs_spec.cuh', 'mali_fs_step.inc', 'spec_instances.inc', os.path.join('..', '..', 'include', 'mali_b200.h')]

EXTRA = os.environ.get('MALI_NVCC_EXTRA', '').split()
NVCC_FLAGS = EXTRA + ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '--fmad=false', '-std=c++20',
              '-shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-ffp-contract=off', '-Xcompiler', '-O2',
              '-cudart', 'static']


def nvcc_path():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError('nvcc not found')


def stale():
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS) or os.path.getmtime(__file__) > t


def build(force=False, verbose=False, lib=None, spec_inc=None):
    """Build libmali_b200.so, or (lib, spec_inc given) a model-specific variant with other kernel instances."""
    lib = LIB if lib is None else lib
    if lib == LIB and not force and not stale():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    extra = ['-DMALI_SPEC_INC="%s"' % spec_inc] if spec_inc else []
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (['-Xptxas', '-v'] if verbose else []) + \
        ['-ccbin', '/usr/bin/g++' if os.path.isfile('/usr/bin/g++') else 'g++'] + \
        ['-o', lib + '.tmp'] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building %s' % os.path.basename(lib))
    os.replace(lib + '.tmp', lib)
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
