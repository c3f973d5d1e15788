"""Grid + metric terms, mirror of the reference `geometry` module (geometry.py:9-182).

Host side: plain float64 numpy in SI base units with the reference's broadcast shapes
(`sig*`: (L,1,1); `dx_j`, `dx_h`: (1,H,1); `dy` scalar; `lat`: (H,1) rad; `long`: (W,) rad; `area`: (H,);
`ptop` Pa; `heightmap`: (H,W) m).  Row 0 is the north-most row, k = 0 the surface layer.
Device side: `device_geom(geom)` uploads the tables once (sigma tables, dx_j, dx_h, heightmap, the
polar-filter multipliers of low_pass.py:61-72 and the FFT twiddles) and keeps them resident; kernels
receive the handle.  A geometry object may be mutated by the caller (e.g. `geom.heightmap[0, 8] = 1000`,
test_geography.py:13); the device copy is refreshed when a fingerprint of the host tables changes.
"""
import ctypes
import hashlib
import math

import numpy as np

from . import _abi, _host, _lib
from .constants import G, Md, R, radius

__all__ = ["Geom", "manabe_sig", "equal_sig", "gen_geometry", "gen_square_geometry", "device_geom", "polar_filter_table",
           "coriolis_parameters", "pressure_from_heightmap"]


class Geom:
    """geometry.py:9-26."""

    def __init__(self, height, width, layers):
        self.height = height
        self.width = width
        self.layers = layers
        self.sige = self.dsig = self.sigb = self.sigt = self.sig = self.dsigv = None
        self.dy = 0.0
        self.lat = 0.0
        self.long = 0.0
        self.dx_j = self.dx_h = None
        self.area = None
        self.ptop = 0.0
        self.heightmap = None
        self.step_options = None      # opt-in terms of the 2.5-D step (dynamics.configure); None = the reference's step
        self._dev = {}

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_dev"] = {}
        return d


def manabe_sig(s):
    """geometry.py:30."""
    return s ** 2 * (3 - 2 * s)


def equal_sig(s):
    """geometry.py:34."""
    return s


def _sigma_tables(geom, layers, sig_func):
    """geometry.py:73-85 / :158-172: sige[k] = sig_func(1 - k/L), k = 0..L."""
    edges = np.asarray([sig_func(1 - k / layers) for k in range(layers + 1)], dtype=np.float64)
    col = lambda a: np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 1, 1)
    geom.sige = col(edges)
    geom.sigt = col(edges[1:])
    geom.sigb = col(edges[:-1])
    geom.dsig = geom.sigb - geom.sigt
    geom.sig = (geom.sigb + geom.sigt) / 2
    geom.dsigv = np.roll(geom.sig, -1, 0) - geom.sig


def gen_geometry(height, width, layers, sig_func=equal_sig,
                 north_edge=90, south_edge=-90, west_edge=-180, east_edge=180):
    """geometry.py:38-151 (the reference's prints are dropped)."""
    geom = Geom(height, width, layers)
    _sigma_tables(geom, layers, sig_func)
    circumference = 2 * radius * math.pi
    dlat = (north_edge - south_edge) / height
    dlong = (east_edge - west_edge) / width
    rows = np.arange(height, dtype=np.float64)
    lat_j = north_edge - (rows + 0.5) * dlat          # cell centres (u, p points)
    lat_h = north_edge - (rows + 1) * dlat            # southern cell edges (v points)
    long_k = west_edge + (np.arange(width, dtype=np.float64) + 0.5) * dlong
    geom.lat = lat_j.reshape(height, -1) * (math.pi / 180.0)
    geom.long = long_k * (math.pi / 180.0)
    geom.dx_j = (np.cos(lat_j * np.pi / 180) * circumference / width).reshape(1, height, 1)
    dx_h = np.cos(lat_h * np.pi / 180) * circumference / width
    geom.dx_h = dx_h.reshape(1, height, 1)
    geom.dy = circumference / 2 / height
    geom.area = (np.roll(dx_h, 1, axis=0) + dx_h) * geom.dy * 0.5   # trapezoids, geometry.py:141
    geom.ptop = 0.0
    geom.heightmap = np.zeros((height, width))
    return geom


def gen_square_geometry(height, width, layers, dx, dy, sig_func=equal_sig):
    """geometry.py:154-182: uniform Cartesian metric."""
    geom = Geom(height, width, layers)
    _sigma_tables(geom, layers, sig_func)
    geom.dx_j = np.full((1, height, 1), _host.scalar(dx))
    geom.dx_h = np.full((1, height, 1), _host.scalar(dx))
    geom.dy = _host.scalar(dy)
    geom.heightmap = np.zeros((height, width))
    return geom


def pressure_from_heightmap(height, sea_level_pressure, sea_level_temp):
    """geometry.py:185-231: isothermal barometric formula p = p0 exp(-G Md h / (R T0)) -- the surface pressure that
    goes with a `geom.heightmap` (host side, once; the reference's prints and abandoned attempts are dropped)."""
    h = np.asarray(_host.magnitude(height), dtype=np.float64)
    p0, t0 = _host.magnitude(sea_level_pressure), _host.magnitude(sea_level_temp)
    return p0 * np.exp((-G * Md * h) / (R * t0))


def polar_filter_table(geom, im=None):
    """Multipliers smmz[j, n] of low_pass.arakawa_1977 (low_pass.py:61-72), shape (H, im/2+1):
    1 for n = 0, min(1, (dx_j/dy) / sin(pi n / im)) for 1 <= n <= im/2."""
    im = geom.width if im is None else im
    drat = _host.scalar(geom.dy) / np.asarray(_host.magnitude(geom.dx_j), dtype=np.float64).reshape(-1, 1)
    bysn = 1 / np.sin(np.pi / im * np.arange(1, im / 2 + 1))
    sm = 1 - bysn / drat
    smmz = 1 - np.maximum(sm, np.zeros_like(sm))
    return np.ascontiguousarray(np.insert(smmz, 0, 1, -1))


def coriolis_parameters(geom):
    """(cp_at_u, cp_at_v) of dynamics.py:89-92 as (H,) arrays: 2 sin(lat) w at the u rows and 2 sin(jph(lat)) w at the
    v rows, w = 2 pi / day; zeros for a geometry without latitudes (gen_square_geometry, geometry.py:174)."""
    w = 2 * math.pi / 86400.0
    lat = np.asarray(_host.magnitude(geom.lat), dtype=np.float64)
    if lat.ndim == 0:
        z = np.zeros(geom.height)
        return z, z.copy()
    lat = lat.reshape(geom.height, -1)
    cp_u = 2 * np.sin(lat) * w
    cp_v = 2 * np.sin((lat + np.roll(lat, -1, -2)) / 2) * w
    return np.ascontiguousarray(cp_u[:, 0]), np.ascontiguousarray(cp_v[:, 0])


def _push_options(geom, dg):
    """Hand geom.step_options to the device geometry (gcm_pe25_set_options)."""
    opt = getattr(geom, "step_options", None)
    if opt is None or not opt.any():
        if dg.options_on:
            _lib.check(_lib.lib().gcm_pe25_set_options(dg.handle, None), "gcm_pe25_set_options")
            dg.options_on = False
        return
    cu, cv = coriolis_parameters(geom) if opt.coriolis else (None, None)
    if cu is not None and dg.rows is not None:      # a latitude band: the values of the rows it stores
        cu, cv = np.ascontiguousarray(cu[dg.rows]), np.ascontiguousarray(cv[dg.rows])
    o = _abi.Pe25Options(int(opt.coriolis), int(opt.limit_q), int(opt.limit_t), float(opt.viscosity),
                         _host.hptr(cu), _host.hptr(cv))
    _lib.check(_lib.lib().gcm_pe25_set_options(dg.handle, ctypes.byref(o)), "gcm_pe25_set_options")
    dg.options_on = True


class DeviceGeom:
    """Owner of one `gcm_geom*` (include/gcm_b200.h)."""

    def __init__(self, handle, H, W, L, row_lo, row_hi, wrap_j):
        self.handle, self.H, self.W, self.L = handle, H, W, L
        self.row_lo, self.row_hi, self.wrap_j = row_lo, row_hi, wrap_j
        self.options_on = False
        self.rows = None              # global row of every stored row (latitude bands only)

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().gcm_geom_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _vec(a, n):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(_host.magnitude(a), dtype=np.float64).reshape(-1), (n,)))


def _fingerprint(geom, band):
    h = hashlib.blake2b(digest_size=16)
    for a in (geom.sige, geom.dx_j, geom.dx_h, geom.heightmap):
        h.update(np.ascontiguousarray(_host.magnitude(a), dtype=np.float64).tobytes())
    h.update(np.asarray([_host.scalar(geom.dy), _host.scalar(geom.ptop)]).tobytes())
    h.update(repr(band).encode())
    h.update(str(_lib.device()).encode())      # tables, streams and events belong to one device
    return h.hexdigest()


def device_geom(geom, band=None):
    """Device-resident tables for `geom`.  band = None: the whole grid, rows periodic in j like np.roll
    (coordinates_3d.py:43-48).  band = (j0, j1, halo_n, halo_s): a latitude band holding global rows
    [j0 - halo_n, j1 + halo_s) (mod H) of which [j0, j1) are owned."""
    key = _fingerprint(geom, band)
    if key in geom._dev:
        return geom._dev[key]
    if band is not None and getattr(geom, "step_options", None) is not None and geom.step_options.any() and (
            band[2] < 2 or band[3] < 2):
        raise ValueError("the opt-in terms (dynamics.configure) reach rows j - 2 ... j + 2: a latitude band needs two "
                         "halo rows on either side")
    H, W, L = geom.height, geom.width, geom.layers
    sig, dsig = _vec(geom.sig, L), _vec(geom.dsig, L)
    sigb, sigt = _vec(geom.sigb, L), _vec(geom.sigt, L)
    dx_j, dx_h = _vec(geom.dx_j, H), _vec(geom.dx_h, H)
    hmap = np.ascontiguousarray(_host.magnitude(geom.heightmap), dtype=np.float64).reshape(H, W)
    smmz = polar_filter_table(geom, W) if W > 1 else None
    zero_v2 = -1
    if band is None:
        Hs, row_lo, row_hi, wrap, zero_v = H, 0, H, 1, H - 1
    else:
        j0, j1, hn, hs = band
        rows = np.arange(j0 - hn, j1 + hs) % H
        Hs, row_lo, row_hi, wrap = len(rows), hn, hn + (j1 - j0), 0
        dx_j, dx_h, hmap = dx_j[rows].copy(), dx_h[rows].copy(), np.ascontiguousarray(hmap[rows])
        smmz = np.ascontiguousarray(smmz[rows]) if smmz is not None else None
        # stored rows that hold the global last row (the wall, dynamics.py:222): among the owned rows and, for the
        # one-exchange schedule, among the halo rows whose predictor the band recomputes
        lo_c, hi_c = max(row_lo - 1, 0), min(row_hi + 2, Hs)
        wall = [int(lo_c + r) for r in np.nonzero(rows[lo_c:hi_c] == H - 1)[0]]
        zero_v = wall[0] if wall else -1
        zero_v2 = wall[1] if len(wall) > 1 else -1
    desc = _abi.GeomDesc(Hs, W, L, wrap, row_lo, row_hi, zero_v, _host.scalar(geom.dy), _host.scalar(geom.ptop),
                         _host.hptr(sig), _host.hptr(dsig), _host.hptr(sigb), _host.hptr(sigt), _host.hptr(dx_j),
                         _host.hptr(dx_h), _host.hptr(hmap), _host.hptr(smmz), zero_v2)
    handle = ctypes.c_void_p()
    _lib.check(_lib.lib().gcm_geom_create(ctypes.byref(desc), ctypes.byref(handle)), "gcm_geom_create")
    obj = DeviceGeom(handle, Hs, W, L, row_lo, row_hi, wrap)
    if len(geom._dev) >= 8:       # a caller that keeps editing the heightmap: drop stale tables nobody else holds
        import sys
        for k in [k for k, v in geom._dev.items() if sys.getrefcount(v) <= 3]:
            del geom._dev[k]      # (dict + loop variable + getrefcount's argument): no Stepper / BandStepper owns it
    geom._dev[key] = obj
    if band is None:
        _push_options(geom, obj)
    else:
        obj.rows = rows
        if obj.row_lo >= 2 and Hs - obj.row_hi >= 2:
            _push_options(geom, obj)
    return obj
