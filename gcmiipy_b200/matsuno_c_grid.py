"""2-D shallow-water Matsuno scheme on the C-grid, mirror of the reference `matsuno_c_grid` module
(matsuno_c_grid.py:15-142).  Uniform doubly periodic grid, fields [j, i].  Only + - * / appear, the
kernels keep the reference's operation order and are compiled without FMA contraction: results are
bit-identical to numpy.
"""
import numpy as np

from . import _host, _lib
from .constants import G


def _op(op, u, v, p, dx):
    fam = _host.Family(u, v, p)
    ts = [(_host.dev(x) if x is not None else None) for x in (u, v, p)]
    ref = next(x for x in ts if x is not None)
    assert ref.dim() == 2, "2-D [j, i] fields"
    H, W = ref.shape
    out = _host.empty((H, W))
    _lib.check(_lib.lib().gcm_sw2d_operator(op, _host.ptr(ts[0]), _host.ptr(ts[1]), _host.ptr(ts[2]), _host.ptr(out),
                                            H, W, _host.scalar(dx), _lib.stream()), "gcm_sw2d_operator")
    return fam.out(out)


def advection_of_velocity_u(u, v, dx):
    """matsuno_c_grid.py:15-51."""
    return _op(0, u, v, None, dx)


def advection_of_velocity_v(u, v, dx):
    """matsuno_c_grid.py:54-80."""
    return _op(1, u, v, None, dx)


def geopotential_gradient_u(p, dx):
    """matsuno_c_grid.py:97-100."""
    return _op(2, None, None, p, dx)


def geopotential_gradient_v(p, dx):
    """matsuno_c_grid.py:103-106."""
    return _op(3, None, None, p, dx)


def advection_of_geopotential(u, v, p, dx):
    """matsuno_c_grid.py:109-118."""
    return _op(4, u, v, p, dx)


def courant_number(p, u, dx, dt):
    """matsuno_c_grid.py:121-122 (host diagnostic)."""
    p, u = (np.asarray(_host.magnitude(x.cpu() if hasattr(x, "cpu") else x), dtype=np.float64) for x in (p, u))
    return (np.max(u) + np.sqrt(np.mean(p) * G)) * _host.scalar(dt) / _host.scalar(dx)


def matsumo_scheme(u, v, p, dx, dt, nsteps=1):
    """matsuno_c_grid.py:125-142: u,v,p -> u',v',p' (predictor + corrector in ONE launch).
    nsteps > 1 advances several steps without leaving the device (grids up to 220 KB of state stay in one
    SM's shared memory for the whole run)."""
    fam = _host.Family(u, v, p)
    tu, tv, tp = (_host.dev(x) for x in (u, v, p))
    assert tu.dim() == 2 and tu.shape == tv.shape == tp.shape
    H, W = tu.shape
    outs = [_host.empty((H, W)) for _ in range(3)]
    need = _lib.lib().gcm_sw2d_workspace_bytes(H, W)
    ws = _host.empty(((need + 7) // 8,))
    _lib.check(_lib.lib().gcm_sw2d_matsuno_step(_host.ptr(tu), _host.ptr(tv), _host.ptr(tp), _host.ptr(outs[0]),
                                                _host.ptr(outs[1]), _host.ptr(outs[2]), H, W, _host.scalar(dx),
                                                _host.scalar(dt), int(nsteps), _host.ptr(ws), need, _lib.stream()),
               "gcm_sw2d_matsuno_step")
    return tuple(fam.out(x, unit) for x, unit in zip(outs, ("meter / second", "meter / second", "meter")))
