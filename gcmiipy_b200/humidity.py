"""Host-side humidity helpers needed by `no_limits_2_5d.gen_initial_conditions`, mirror of the
reference `humidity` module (humidity.py:4-31).  Initialisation only; never on the device path."""
import numpy as np

from .constants import Rd, Rv


def manabe_rh(geom):
    """humidity.py:4-7."""
    return 0.77 * (geom.sig - 0.02) / (1 - 0.02)


def saturation_vapor_pressure(tt):
    """Buck equation, Pa (humidity.py:10-14)."""
    t = np.asarray(tt, dtype=np.float64) - 273.15
    return 0.61121 * 1e3 * np.exp((18.678 - t / 234.5) * (t / (257.14 + t)))


def rh_to_mmr(rh, tp, tt):
    """humidity.py:27-37."""
    e = rh * saturation_vapor_pressure(tt)
    w = e * Rd / (Rv * (tp - e))
    return w / (w + 1)
