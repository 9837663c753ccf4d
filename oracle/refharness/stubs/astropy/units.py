"""Tiny exact-rational unit system: just enough of `astropy.units` for the
reference's setup code (shim 1).

Scale factors are kept as `fractions.Fraction` so that conversions such as
cm^-3 -> m^-3 are the exact decimal 1e6 rather than `0.01**-3`.
"""
from fractions import Fraction
import sys
import numpy as np


class Unit:
    __array_ufunc__ = None  # make ndarray binary operators defer to us

    def __init__(self, scale=Fraction(1), dims=None, name=None):
        self.scale = Fraction(scale)
        self.dims = {k: v for k, v in (dims or {}).items() if v != 0}
        self.name = name

    def _combine(self, other, sign):
        dims = dict(self.dims)
        for k, v in other.dims.items():
            dims[k] = dims.get(k, 0) + sign * v
        scale = self.scale * other.scale if sign > 0 else self.scale / other.scale
        return Unit(scale, dims)

    def __mul__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, +1)
        return Quantity(np.asarray(other, dtype=float), self)

    def __rmul__(self, other):
        return Quantity(np.asarray(other, dtype=float), self)

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return self._combine(other, -1)
        return NotImplemented

    def __rtruediv__(self, other):
        return Quantity(np.asarray(other, dtype=float), self ** -1)

    def __pow__(self, p):
        p = int(p)
        return Unit(self.scale ** p, {k: v * p for k, v in self.dims.items()})

    def is_equivalent(self, other):
        return self.dims == other.dims

    def __rlshift__(self, value):
        if isinstance(value, Quantity):
            return value.to(self)
        return Quantity(np.array(value, dtype=float), self)

    def __eq__(self, other):
        return isinstance(other, Unit) and self.dims == other.dims and self.scale == other.scale

    def __hash__(self):
        return hash((self.scale, tuple(sorted(self.dims.items()))))

    def __repr__(self):
        return 'Unit(%s, %s)' % (self.scale, self.dims)


class Quantity(np.ndarray):
    def __new__(cls, value, unit=None, dtype=None, copy=True):
        arr = np.array(value, dtype=float, copy=True if copy else None)
        obj = arr.view(cls)
        obj.unit = unit if unit is not None else one
        return obj

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.unit = getattr(obj, 'unit', one)

    @property
    def value(self):
        return self.view(np.ndarray)

    def to(self, unit, equivalencies=None):
        if not self.unit.is_equivalent(unit):
            raise ValueError('stub units: %r not convertible to %r' % (self.unit, unit))
        factor = self.unit.scale / unit.scale
        if factor == 1:
            out = np.array(self.view(np.ndarray), dtype=float, copy=True)
        else:
            out = self.view(np.ndarray) * float(factor)
        return Quantity(out, unit)

    def __lshift__(self, unit):
        if isinstance(unit, Unit):
            return self.to(unit)
        return super().__lshift__(unit)


def quantity_input(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]

    def deco(fn):
        return fn
    return deco


def spectral_density(wav):
    raise NotImplementedError('stub astropy: spectral_density is not available')


one = Unit(1, {}, 'one')
m = Unit(1, {'m': 1}, 'm')
s = Unit(1, {'s': 1}, 's')
kg = Unit(1, {'kg': 1}, 'kg')
K = Unit(1, {'K': 1}, 'K')
sr = Unit(1, {'sr': 1}, 'sr')
g = Unit(Fraction(1, 1000), {'kg': 1}, 'g')
cm = Unit(Fraction(1, 100), {'m': 1}, 'cm')
km = Unit(1000, {'m': 1}, 'km')
nm = Unit(Fraction(1, 10**9), {'m': 1}, 'nm')
Hz = Unit(1, {'s': -1}, 'Hz')
J = Unit(1, {'kg': 1, 'm': 2, 's': -2}, 'J')

quantity = sys.modules[__name__]
