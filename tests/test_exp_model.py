"""CPU check of the exp() algorithm the CUDA kernels use (csrc/mali_kernels.cuh: exp_m) against libm's exp,
which is what the reference's numba code calls in formal_solver.py:41.  The algorithm is re-stated here with
exact-rational arithmetic (every operation rounded once to nearest-even, fused multiply-adds rounded once), and
the 2^(k/128) table is read from the generated csrc/exp_table.inc -- so this pins both the operation sequence
and the table bit for bit.  The GPU twin is tests/test_gpu_parity.py::test_exp_hook_bitwise.
"""
import math
import os
import re
import struct
from fractions import Fraction

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, 'lightspinner_b200', 'csrc', 'exp_table.inc')


def _bits(x):
    return struct.unpack('<Q', struct.pack('<d', x))[0]


def _dbl(b):
    return struct.unpack('<d', struct.pack('<Q', b & 0xFFFFFFFFFFFFFFFF))[0]


def _rn(fr):
    return float(fr)   # Fraction -> float is correctly rounded (nearest even)


def fma(a, b, c):
    return _rn(Fraction(a) * Fraction(b) + Fraction(c))


def load_table():
    tab = []
    for m in re.finditer(r'\{0x([0-9a-f]+)ull, 0x([0-9a-f]+)ull\}', open(INC).read()):
        tab.append((int(m.group(1), 16), int(m.group(2), 16)))
    assert len(tab) == 128
    return tab


def exp_model(x, tab):
    InvLn2N, Shift = float.fromhex('0x1.71547652b82fep+7'), float.fromhex('0x1.8p+52')
    NegLn2hiN, NegLn2loN = float.fromhex('-0x1.62e42fefa0000p-8'), float.fromhex('-0x1.cf79abc9e3b3ap-47')
    C2, C3 = float.fromhex('0x1.ffffffffffdbdp-2'), float.fromhex('0x1.555555555543cp-3')
    C4, C5 = float.fromhex('0x1.55555cf172b91p-5'), float.fromhex('0x1.1111167a4d017p-7')
    kd = fma(x, InvLn2N, Shift)
    ki = _bits(kd)
    kd = kd - Shift
    r = fma(kd, NegLn2hiN, x)
    r = fma(kd, NegLn2loN, r)
    tb, hb = tab[ki & 127]
    tail = _dbl(tb)
    scale = _dbl(hb + (ki << 45))
    t1 = fma(r, C3, C2)
    s = r + tail
    r2 = r * r
    t2 = fma(r, C5, C4)
    s2 = fma(t1, r2, s)
    r4 = r2 * r2
    tmp = fma(r4, t2, s2)
    return fma(scale, tmp, scale)


def test_table_is_2_pow_k_over_128():
    tab = load_table()
    for k, (tb, hb) in enumerate(tab):
        H = _dbl(hb + (k << 45))
        assert abs(H / 2.0 ** (k / 128.0) - 1.0) < 3e-16
        assert abs(_dbl(tb)) < 2.0 ** -53


def test_exp_model_equals_libm_bitwise():
    tab = load_table()
    rng = np.random.default_rng(2026)
    xs = np.concatenate([-np.exp(rng.uniform(np.log(5e-4), np.log(50.0), 6000)),      # the range w2 uses
                         -rng.uniform(5e-4, 50.0, 2000), [-5e-4, -50.0, -1.0, -0.6931471805599453, -1e-3]])
    bad = [x for x in xs if exp_model(float(x), tab) != math.exp(float(x))]
    assert not bad, 'exp model differs from libm for %d of %d arguments, e.g. %r' % (len(bad), len(xs), bad[:3])
