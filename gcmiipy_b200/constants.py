"""Physical constants in SI base units, mirror of the reference `constants` module (constants.py:16-78).

The reference wraps every value in a pint Quantity; here units are stripped at the boundary (_host.py)
and the kernels are unit-free, so the constants are plain floats.  `unit_roll` is np.roll
(constants.py:85); on the device every shift is index arithmetic inside the kernels.
"""
import numpy as np

R = 8.3145                     # J / (K mol)
Md = 28.97e-3                  # kg / mol
Rd = 287.0                     # J / (K kg)      constants.py:16
rd = 1.275                     # kg / m^3
Cp = 1004.0                    # J / (K kg)      constants.py:22
Cg = 1.13e6                    # J / (K m^3)
kappa = Rd / Cp                # constants.py:28
P0 = 100000.0                  # Pa              constants.py:31
standard_pressure = 101325.0   # Pa
standard_temperature = 273.16  # K
G = 9.8                        # m / s^2         constants.py:45
radius = 6.3781e6              # m               constants.py:48
mu_air = 18.5 * 1e-6           # Pa s            constants.py:51
Rv = 461.0                     # J / (K kg)      constants.py:78
x_dim, y_dim, z_dim = -1, -2, -3


def unit_roll(a, shift, axis=None):
    """constants.py:85-89 (host arrays)."""
    return np.roll(a, shift, axis=axis)


def get_total_variation(q):
    """constants.py:105-108 (host diagnostic)."""
    q = np.asarray(getattr(q, "magnitude", q), dtype=np.float64)
    return np.sum(np.abs(q - np.roll(q, -1, 0)))


def courant_number(p, u, dx, dt):
    """constants.py:111-112 (host diagnostic)."""
    from . import _host
    p, u = (np.asarray(_host.magnitude(x), dtype=np.float64) for x in (p, u))
    return (np.max(u) + np.sqrt(np.mean(p) * G)) * _host.scalar(dt) / _host.scalar(dx)
