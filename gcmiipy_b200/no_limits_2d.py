"""2-D primitive equations (p, u, v, theta; q passive) Matsuno step on the C-grid, mirror of the
reference `no_limits_2d` module (no_limits_2d.py:21-131).  Uniform doubly periodic grid, scalar dx."""
import ctypes

import torch

from . import _host, _lib
from .dynamics import _struct, calc_pu, calc_pv, un_pu, un_pv  # noqa: F401  (same helpers, :21-38)

_UNITS = ("pascal", "meter / second", "meter / second", "kelvin", "dimensionless")


def _op(op, p, u, v, t, dx):
    fam = _host.Family(p, u, v, t)
    ts = [(_host.dev(x) if x is not None else None) for x in (p, u, v, t)]
    H, W = ts[0].shape
    o0, o1 = _host.empty((H, W)), _host.empty((H, W))
    _lib.check(_lib.lib().gcm_pe2d_operator(op, *[_host.ptr(x) for x in ts], _host.ptr(o0), _host.ptr(o1), H, W,
                                            _host.scalar(dx), _lib.stream()), "gcm_pe2d_operator")
    return fam.out(o0), fam.out(o1)


def advec_m(p, u, v, dx):
    """no_limits_2d.py:47-73 -> (dut, dvt)."""
    return _op(0, p, u, v, None, dx)


def pgf(p, t, dx):
    """no_limits_2d.py:76-89 -> (pgfu, pgfv)."""
    return _op(1, p, None, None, t, dx)


def half_timestep(p, u, v, t, q, sp, su, sv, st, sq, dt, dx):
    """no_limits_2d.py:104-126."""
    fam = _host.Family(p, u, v, t, q, sp, su, sv, st, sq)
    base = [_host.dev(x) for x in (p, u, v, t, q)]
    star = [_host.dev(x) for x in (sp, su, sv, st, sq)]
    H, W = base[0].shape
    out = [torch.empty_like(x) for x in base]
    sb, ss, so = _struct(base), _struct(star), _struct(out)
    _lib.check(_lib.lib().gcm_pe2d_half_step(ctypes.byref(sb), ctypes.byref(ss), ctypes.byref(so), H, W, _host.scalar(dt),
                                             _host.scalar(dx), _lib.stream()), "gcm_pe2d_half_step")
    return tuple(fam.out(x, unit) for x, unit in zip(out, _UNITS))


def matsuno_timestep(p, u, v, t, q, dt, dx, nsteps=1):
    """no_limits_2d.py:129-131; nsteps > 1 keeps the state on the device between steps."""
    fam = _host.Family(p, u, v, t, q)
    base = [_host.dev(x) for x in (p, u, v, t, q)]
    H, W = base[0].shape
    out = [torch.empty_like(x) for x in base]
    need = _lib.lib().gcm_pe2d_workspace_bytes(H, W)
    ws = _host.empty(((need + 7) // 8,))
    sb, so = _struct(base), _struct(out)
    _lib.check(_lib.lib().gcm_pe2d_matsuno_step(ctypes.byref(sb), ctypes.byref(so), H, W, _host.scalar(dt),
                                                _host.scalar(dx), int(nsteps), _host.ptr(ws), need, _lib.stream()),
               "gcm_pe2d_matsuno_step")
    return tuple(fam.out(x, unit) for x, unit in zip(out, _UNITS))
