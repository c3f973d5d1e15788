"""Grey-radiation column physics, mirror of the reference `grey_solar` module for the functions on the path of
`no_limits_2_5d.solar_timestep` (grey_solar.py:40-68 zenith angle, :323-333 transmittances, :358-563
basic_grey_radiation): same names and positional signatures; the column arithmetic runs in `csrc/physics.cu` (one
thread per column) behind `gcm_grey_radiation` / `gcm_solar_timestep` (include/gcm_b200.h).

Times (`utc`) are seconds, or Quantities; angles radians.  The experimental variants of the reference file
(`grey_solar`, `grey_radiation`, `basic_3_gas_absorbance`: unreachable from the model driver) are out of scope.
"""
import math

import numpy as np
import torch

from . import _host, _lib
from .geometry import device_geom

sb_constant = 5.67e-8              # W m-2 K-4   constants.py:71
solar_constant = 1.3608 * 1000.0   # W m-2       constants.py:59


def solar_zenith_angle(latitude, hour_angle, declination):
    """grey_solar.py:40-46: cosine of the solar zenith angle (host arrays)."""
    return np.sin(latitude) * np.sin(declination) + np.cos(latitude) * np.cos(declination) * np.cos(hour_angle)


def hour_angle(time):
    """grey_solar.py:51: negative because the sun moves west; radians."""
    return _host.scalar(time) / (-24 * 3600.0) * 360 * (math.pi / 180.0)


def zenith_angle(longs, lats, time, geom):
    """grey_solar.py:49-68: max(cos zenith, 0) on the [H, W] grid at model time `time` (declination 0)."""
    longs = np.asarray(_host.magnitude(longs), dtype=np.float64)
    lats = np.asarray(_host.magnitude(lats), dtype=np.float64).reshape(geom.height, -1)
    point_angle = np.tile(longs, (geom.height, 1)) + hour_angle(time)
    return np.maximum(solar_zenith_angle(lats, point_angle, 0.0), 0)


def basic_grey_transmittances(t_lw, t_sw, geom):
    """grey_solar.py:323-333 (AD 2.35): per-layer long-wave / short-wave transmittance, shape (L, 1, 1)."""
    dsig = np.asarray(_host.magnitude(geom.dsig), dtype=np.float64)
    e_n = 1 - t_lw ** dsig
    e_n_sw = 1 - t_sw ** dsig
    return 1 - e_n, 1 - e_n_sw


def _tables(geom, dg):
    """sin / cos of the latitudes and the longitudes as device tables, cached on the device geometry."""
    tabs = getattr(dg, "_solar_tabs", None)
    if tabs is None:
        lat = np.asarray(_host.magnitude(geom.lat), dtype=np.float64).reshape(-1)
        lon = np.asarray(_host.magnitude(geom.long), dtype=np.float64).reshape(-1)
        if getattr(dg, "rows", None) is not None:      # a latitude band: the stored rows' latitudes (halo rows included)
            lat = lat[np.asarray(dg.rows)]
        mk = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(_lib.device())
        tabs = (mk(np.sin(lat)), mk(np.cos(lat)), mk(lon))
        dg._solar_tabs = tabs
    return tabs


def _ground_temperature(g):
    """g: GroundVars (no_limits_2_5d.py:143: gt, gw, snow, ice) or the ground temperature array / tensor itself (a torch
    tensor has a `.gt` METHOD, so the namedtuple is recognised by its fields)."""
    return g.gt if "gt" in getattr(g, "_fields", ()) else g


def _layer_arrays(t_lw, t_sw, geom):
    lw, sw = basic_grey_transmittances(_host.scalar(t_lw), _host.scalar(t_sw), geom)
    return np.ascontiguousarray(lw.reshape(-1)), np.ascontiguousarray(sw.reshape(-1))


def basic_grey_radiation(p, tp, tt, g, t_lw, t_sw, albedo, utc, geom):
    """grey_solar.py:358-563: the basic grey atmosphere of AD 2.7 -> (dT/dt [L, H, W] in K/s, d(ground T)/dt [H, W]).
    p: surface pressure, tp: layer pressures (unused, as in the reference), tt: true temperature, g: GroundVars (or the
    ground temperature itself)."""
    gt = _ground_temperature(g)
    fam = _host.Family(p, tt, gt)
    dg = device_geom(geom)
    dp, dtt, dgt = (_host.dev(x).contiguous() for x in (p, tt, gt))
    L, H, W = dg.L, dg.H, dg.W
    assert tuple(dtt.shape) == (L, H, W) and tuple(dp.shape) == (H, W) and tuple(dgt.shape) == (H, W)
    lw, sw = _layer_arrays(t_lw, t_sw, geom)
    sinlat, coslat, lon = _tables(geom, dg)
    dTdt, dtg = _host.empty((L, H, W)), _host.empty((H, W))
    _lib.check(_lib.lib().gcm_grey_radiation(dg.handle, _host.ptr(dp), _host.ptr(dtt), _host.ptr(dgt), _host.hptr(lw),
                                            _host.hptr(sw), _host.scalar(albedo), _host.ptr(sinlat), _host.ptr(coslat),
                                            _host.ptr(lon), hour_angle(utc), _host.ptr(dTdt), _host.ptr(dtg),
                                            _lib.stream()), "gcm_grey_radiation")
    return fam.out(dTdt, "kelvin / second"), fam.out(dtg, "kelvin / second")


def solar_timestep(t, p, g, dt, utc, geom, t_lw=0.1, t_sw=0.9, albedo=0.3, dg=None):
    """no_limits_2_5d.solar_timestep (no_limits_2_5d.py:66-75) in one launch: theta -> T, radiation, explicit update
    of the air and ground temperatures over dt, T -> theta.  Returns (theta_n, ground temperature_n).
    dg: the device geometry of a latitude band (bands.BandStepper.dg) -- the columns are independent, so a band runs
    the same kernel on its stored rows (arrays shaped like the band)."""
    gt = _ground_temperature(g)
    fam = _host.Family(t, p, gt)
    dg = dg if dg is not None else device_geom(geom)
    dt_, dp, dgt = (_host.dev(x).contiguous() for x in (t, p, gt))
    L, H, W = dg.L, dg.H, dg.W
    assert tuple(dt_.shape) == (L, H, W) and tuple(dp.shape) == (H, W) and tuple(dgt.shape) == (H, W)
    lw, sw = _layer_arrays(t_lw, t_sw, geom)
    sinlat, coslat, lon = _tables(geom, dg)
    t_n, gt_n = _host.empty((L, H, W)), _host.empty((H, W))
    _lib.check(_lib.lib().gcm_solar_timestep(dg.handle, _host.ptr(dp), _host.ptr(dt_), _host.ptr(dgt), _host.hptr(lw),
                                            _host.hptr(sw), _host.scalar(albedo), _host.ptr(sinlat), _host.ptr(coslat),
                                            _host.ptr(lon), hour_angle(utc), _host.scalar(dt), _host.ptr(t_n),
                                            _host.ptr(gt_n), _lib.stream()), "gcm_solar_timestep")
    return fam.out(t_n, "kelvin"), fam.out(gt_n, "kelvin")
