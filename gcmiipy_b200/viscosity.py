"""5-point Laplacian and mu * lap(u), mirror of the reference `viscosity` module (viscosity.py:12-25)."""
from . import _host, _lib


def _lap(q, dx, mu, apply_mu):
    fam = _host.Family(q)
    t = _host.dev(q)
    assert t.dim() == 2, "viscosity operators take a 2-D [j, i] field"
    out = _host.empty(t.shape)
    _lib.check(_lib.lib().gcm_laplacian5(_host.ptr(t), _host.ptr(out), t.shape[0], t.shape[1], _host.scalar(dx),
                                         float(mu), apply_mu, _lib.stream()), "gcm_laplacian5")
    return fam.out(out)


def finite_laplacian_2d(q, dx):
    """(q[j+1] + q[j-1] + q[i+1] + q[i-1] - 4 q) / dx^2, doubly periodic   (viscosity.py:12-19)."""
    return _lap(q, dx, 1.0, 0)


def incompressible_viscosity_2d(u, mu, dx):
    """mu * lap(u)   (viscosity.py:22-25)."""
    return _lap(u, dx, _host.scalar(mu), 1)
