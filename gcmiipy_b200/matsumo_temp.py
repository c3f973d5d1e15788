"""2-D Matsuno shallow water + temperature + viscosity, mirror of the reference `matsumo_temp` module
(matsumo_temp.py:13-99), the only in-reference consumer of `viscosity`.
Keeps the reference quirk that the v equation is damped with the Laplacian of u (:75, :91)."""
from . import _host, _lib
from .constants import Cp, Rd, mu_air, standard_pressure, standard_temperature
from .matsuno_c_grid import *  # noqa: F401,F403  (the reference imports the five operators the same way, :8-9)


def density_from(p, t):
    """matsumo_temp.py:13-16 (host helper)."""
    import numpy as np
    p, t = (np.asarray(_host.magnitude(x), dtype=float) for x in (p, t))
    temp = t / ((100000.0 / p) ** (Rd / Cp))
    return p / (Rd * temp)


def gen_initial_conditions(side_len):
    """matsumo_temp.py:100-105: u = v = 0, p = standard pressure, t = standard temperature."""
    import numpy as np
    shape = (side_len, side_len)
    return np.zeros(shape), np.zeros(shape), np.full(shape, standard_pressure), np.full(shape, standard_temperature)


def matsumo_scheme(u, v, p, t, dx, dt, nsteps=1, mu=mu_air):
    """matsumo_temp.py:66-99 -> (u', v', p', t')."""
    fam = _host.Family(u, v, p, t)
    ts = [_host.dev(x) for x in (u, v, p, t)]
    H, W = ts[0].shape
    outs = [_host.empty((H, W)) for _ in range(4)]
    need = _lib.lib().gcm_swt2d_workspace_bytes(H, W)
    ws = _host.empty(((need + 7) // 8,))
    _lib.check(_lib.lib().gcm_swt2d_matsuno_step(*[_host.ptr(x) for x in ts], *[_host.ptr(x) for x in outs], H, W,
                                                 _host.scalar(dx), _host.scalar(dt), _host.scalar(mu), int(nsteps),
                                                 _host.ptr(ws), need, _lib.stream()), "gcm_swt2d_matsuno_step")
    return tuple(fam.out(x) for x in outs)
