"""Minimal SI-magnitude stand-in for the third-party `pint` package.

TEST INFRASTRUCTURE ONLY.  The reference (marthinwurer/gcmiipy) wraps every
array in a pint Quantity (constants.py:2-5); pint is not installable in this
environment (no network).  pint performs no arithmetic of its own beyond
unit-scale multiplications, so this stand-in stores the SI-base magnitude in an
ndarray subclass and ignores dimensions.  It exists only so that
`oracle/make_golden.py` can import the UNMODIFIED reference from
/root/reference and snapshot its outputs.  Nothing in the product imports it.
"""
import math

import numpy as np

__version__ = "0.0-si-standin"

# scale factor to SI base units for every unit name the reference touches
_SCALE = {
    "m": 1.0, "meter": 1.0, "km": 1e3, "s": 1.0, "second": 1.0, "kg": 1.0, "g": 1e-3, "gram": 1e-3,
    "K": 1.0, "kelvin": 1.0, "mol": 1.0, "J": 1.0, "W": 1.0, "kW": 1e3, "Pa": 1.0, "hPa": 1e2,
    "kPa": 1e3, "uPa": 1e-6, "day": 86400.0, "days": 86400.0, "hour": 3600.0, "hours": 3600.0,
    "minute": 60.0, "minutes": 60.0, "degrees": math.pi / 180.0, "degree": math.pi / 180.0,
    "radian": 1.0, "radians": 1.0, "dimensionless": 1.0, "N": 1.0, "um": 1e-6, "cm": 1e-2, "mm": 1e-3,
    "celsius": 1.0, "degC": 1.0,
}


class Unit:
    """A unit is just its scale to SI base; `is_celsius` marks the one offset unit."""
    __array_ufunc__ = None  # make ndarray defer to our __rmul__/__rtruediv__
    __array_priority__ = 1000

    def __init__(self, scale=1.0, is_celsius=False):
        self.scale = float(scale)
        self.is_celsius = is_celsius

    def __mul__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale * other.scale)
        return Quantity(np.asarray(other, dtype=float) * self.scale)

    __rmul__ = __mul__

    def __truediv__(self, other):
        if isinstance(other, Unit):
            return Unit(self.scale / other.scale)
        return Quantity(self.scale / np.asarray(other, dtype=float))

    def __rtruediv__(self, other):
        return Quantity(np.asarray(other, dtype=float) / self.scale)

    def __pow__(self, e):
        return Unit(self.scale ** float(e))

    def __eq__(self, other):  # dimensions are not tracked: every unit compares equal
        return isinstance(other, Unit)

    def __hash__(self):
        return 0

    def __repr__(self):
        return "<si>"


class Quantity(np.ndarray):
    """ndarray holding SI-base magnitudes; the pint attributes the reference uses."""
    __array_priority__ = 100

    def __new__(cls, value, unit=None):
        arr = np.asarray(value, dtype=float)
        if unit is not None:
            if unit.is_celsius:
                arr = arr + 273.15
            else:
                arr = arr * unit.scale
        return arr.view(cls)

    @property
    def m(self):
        return self.view(np.ndarray)

    magnitude = m

    @property
    def u(self):
        return Unit(1.0)

    units = u

    def to_base_units(self):
        return self

    def to(self, unit):
        if unit.is_celsius:
            return Quantity(self.view(np.ndarray) - 273.15)
        return Quantity(self.view(np.ndarray) / unit.scale)

    def __format__(self, spec):
        if self.ndim == 0:
            return format(float(self), spec)
        return str(self)


class UnitRegistry:
    Quantity = Quantity

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name not in _SCALE:
            raise AttributeError("pint stand-in: unknown unit %r" % name)
        return Unit(_SCALE[name], is_celsius=name in ("celsius", "degC"))
