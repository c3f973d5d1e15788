"""Donor-cell flux, slope ratio and van Leer limiter, mirror of the reference `flux_limiter`
(flux_limiter.py:10-32).  1-D periodic rows; an [nrows, n] array is treated as independent rows."""
from . import _host, _lib


def _rows(t):
    return (1, t.shape[0]) if t.dim() == 1 else (int(t.numel() // t.shape[-1]), t.shape[-1])


def van_leer(r):
    """phi(r) = (r + |r|) / (1 + |r|)   (flux_limiter.py:10-11)."""
    fam = _host.Family(r)
    t = _host.dev(r)
    out = _host.empty(t.shape)
    _lib.check(_lib.lib().gcm_fl_van_leer(_host.ptr(t), _host.ptr(out), t.numel(), _lib.stream()), "gcm_fl_van_leer")
    if t.dim() == 0 and not fam.torch:
        return float(out.item())
    return fam.out(out)


def calc_r(q):
    """r = (q_i - q_{i-1}) / (q_{i+1} - q_i), 0 where the denominator is 0   (flux_limiter.py:14-20)."""
    fam = _host.Family(q)
    t = _host.dev(q)
    nrows, n = _rows(t)
    out = _host.empty(t.shape)
    _lib.check(_lib.lib().gcm_fl_calc_r(_host.ptr(t), _host.ptr(out), nrows, n, _lib.stream()), "gcm_fl_calc_r")
    return fam.out(out)


def donor_cell_flux(q, u):
    """Upwind edge flux u * (q_i if u > 0 else q_{i+1})   (flux_limiter.py:23-27)."""
    fam = _host.Family(q, u)
    tq, tu = _host.dev(q), _host.dev(u)
    nrows, n = _rows(tq)
    out = _host.empty(tq.shape)
    _lib.check(_lib.lib().gcm_fl_donor_cell_flux(_host.ptr(tq), _host.ptr(tu), _host.ptr(out), nrows, n, _lib.stream()),
               "gcm_fl_donor_cell_flux")
    return fam.out(out)


def donor_cell_advection(q, u, dx, dt, nsteps=1):
    """q + (F_{i-1} - F_i) dt / dx   (flux_limiter.py:30-32); nsteps > 1 keeps the field on the device."""
    fam = _host.Family(q, u)
    tq, tu = _host.dev(q), _host.dev(u)
    nrows, n = _rows(tq)
    out = _host.empty(tq.shape)
    tmp = _host.empty(tq.shape) if nsteps > 1 else None
    _lib.check(_lib.lib().gcm_fl_donor_cell_advection(_host.ptr(tq), _host.ptr(tu), _host.ptr(out), nrows, n,
                                                      _host.scalar(dx), _host.scalar(dt), nsteps, _host.ptr(tmp),
                                                      _lib.stream()), "gcm_fl_donor_cell_advection")
    return fam.out(out)
