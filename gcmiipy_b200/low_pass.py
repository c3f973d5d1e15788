"""Polar filter, mirror of the reference `low_pass` module (low_pass.py:14-78).

`arakawa_1977`: per latitude row rfft_i -> x smmz[j, n] -> irfft_i.  On the device one CTA filters two
layers of one row at a time with a shared-memory mixed-radix Stockham FFT (csrc/fft_rows.h); the
multiplier table smmz is precomputed on the host with the reference's own formula and kept resident.
"""
import numpy as np

from . import _host, _lib
from .geometry import device_geom


def _filter(q, geom, table):
    fam = _host.Family(q)
    t = _host.dev(q)
    H, W = geom.height, geom.width
    if t.shape[-1] != W or t.shape[-2] != H:
        raise ValueError("field rows/columns %s do not match the geometry (%d, %d)" % (tuple(t.shape[-2:]), H, W))
    if W == 1:
        return q                                      # low_pass.py:58-59
    if W % 2:
        raise ValueError("the polar filter needs an even number of columns (low_pass.py:57)")
    dg = device_geom(geom)
    nlayers = int(t.numel() // (H * W))
    out = _host.empty(t.shape if t.dim() == 3 else (1,) * (3 - t.dim()) + tuple(t.shape))
    dtab = _host.dev(table) if table is not None else None
    _lib.check(_lib.lib().gcm_polar_filter(dg.handle, _host.ptr(t), _host.ptr(out), nlayers, _host.ptr(dtab),
                                           _lib.stream()), "gcm_polar_filter")
    return fam.out(out)


def arakawa_1977(q, geom):
    """low_pass.py:41-78.  As in the reference a 2-D input comes back with shape (1, H, W) (dx_j is (1,H,1))."""
    return _filter(q, geom, None)


def avrx(q, geom):
    """low_pass.py:14-38: hard spectral cut-off, wavenumbers with rfftfreq(W, dx_j) * dy > 0.5 are zeroed.
    2-D input only; returns (1, H, W) like the reference's broadcast."""
    if np.ndim(_host.magnitude(q)) != 2:
        raise ValueError("avrx takes a 2-D [j, i] field (low_pass.py:15)")
    im = geom.width
    dx_j = np.asarray(_host.magnitude(geom.dx_j), dtype=np.float64).reshape(-1, 1)
    ratios = np.fft.rfftfreq(im, dx_j) * _host.scalar(geom.dy)
    table = np.zeros_like(ratios)
    table[ratios <= 0.5] = 1
    return _filter(q, geom, np.ascontiguousarray(table))
