"""Hydrostatic geopotential column integration, mirror of the reference `phi_port` module
(phi_port.py:5-136), the line-by-line port of the GCM II Fortran PGF vertical differencing."""
import numpy as np
import torch

from . import _host, _lib
from .constants import kappa
from .geometry import device_geom


def PGF(T, P, geom):
    """phi_port.py:5-113.  T[W, H, L] and P[W, H] are the transposed views the reference is called with
    (test_phi_port.py:24-25); returns PHI[W, H, L].  Bug-for-bug: only I = 0 is computed for every J
    (IMAX = 1, :50-54), the rest of PHI is zero; theta-bar is the arithmetic mean (:78)."""
    fam = _host.Family(T, P)
    dg = device_geom(geom)
    if isinstance(T, torch.Tensor):
        t, p = T.permute(2, 1, 0), P.permute(1, 0)
    else:
        t, p = np.transpose(_host.magnitude(T)), np.transpose(_host.magnitude(P))
    t, p = _host.dev(t), _host.dev(p)
    if tuple(t.shape) != (dg.L, dg.H, dg.W) or tuple(p.shape) != (dg.H, dg.W):
        raise ValueError("T must be [W, H, L] and P [W, H] of the geometry")
    phi = _host.empty(t.shape)
    _lib.check(_lib.lib().gcm_phi_port_pgf(dg.handle, _host.ptr(p), _host.ptr(t), _host.ptr(phi), _lib.stream()),
               "gcm_phi_port_pgf")
    out = phi.permute(2, 1, 0).contiguous()
    return fam.out(out)


def EXPBYK(X):
    """phi_port.py:116-117 (host scalar helper)."""
    return X ** kappa


def THBAR(X, Y):
    """phi_port.py:120-136 (host scalar helper; unused by PGF, which takes the arithmetic mean)."""
    x = X / Y
    return X * (np.log(x) / (x - 1))
