"""Driver of the 2.5-D model, mirror of the reference `no_limits_2_5d` module
(no_limits_2_5d.py:35-60 calc_energy, :79-94 full_timestep, :146-168 gen_initial_conditions,
:220-236 run_model).  The column physics after the dynamics step is unreachable in the reference
(`return` at :94) and is out of scope here too.
"""
from collections import defaultdict, namedtuple

import numpy as np
import torch

from . import _host, _lib, geometry
from .constants import P0, kappa
from .dynamics import *  # noqa: F401,F403  (the reference re-exports dynamics the same way, :29)
from .dynamics import Stepper, _struct, matsuno_timestep
from .geometry import device_geom, gen_geometry
from .humidity import manabe_rh, rh_to_mmr

GroundVars = namedtuple("GroundVars", ("gt", "gw", "snow", "ice"))
STATS = defaultdict(list)

height = 24
width = 36
layers = 9


def gen_initial_conditions(geom):
    """no_limits_2_5d.py:146-168 (host side, once): p = 1e5 Pa - ptop, u = 1, v = 0, T = 360 K -> theta,
    q = max(3e-6, Manabe RH mixing ratio); ground variables ride along untouched."""
    full = (geom.layers, geom.height, geom.width)
    surface = (geom.height, geom.width)
    ptop = _host.scalar(geom.ptop)
    p = np.full(surface, 1) * 100000.0 - ptop
    u = np.full(full, 1) * 1.0
    v = np.full(full, 1) * .0
    tt = np.full(full, 1) * 360.0
    tp = p * geom.sig + ptop
    t = tt * ((P0 / tp) ** kappa)
    q = np.maximum(np.full(full, 1) * 0.000003, rh_to_mmr(manabe_rh(geom), tp, tt))
    g = GroundVars(np.full(surface, 1) * 360.0, np.zeros(surface), np.zeros(surface), np.zeros(surface))
    return p, u, v, t, q, g


def _area_by_i(geom):
    """The reference multiplies an (L, H, W) array by geom.area of shape (H,) (no_limits_2_5d.py:49), which
    numpy broadcasts along i: valid only when H == W or H == 1."""
    area = np.asarray(_host.magnitude(geom.area), dtype=np.float64).reshape(-1)
    if area.shape[0] not in (1, geom.width):
        raise ValueError("operands could not be broadcast together with shapes (%d,%d,%d) (%d,)"
                         % (geom.layers, geom.height, geom.width, area.shape[0]))
    return np.ascontiguousarray(np.broadcast_to(area, (geom.width,)))


def calc_energy(p, u, v, t, q, g, geom):
    """no_limits_2_5d.py:35-60 -> (ke, cpT, geopotential, total) in J, one fused reduction kernel."""
    import ctypes
    dg = device_geom(geom)
    ts = [_host.dev(x) for x in (p, u, v, t, q)]
    area = _host.dev(_area_by_i(geom))
    out = _host.empty((3,))
    s = _struct(ts)
    _lib.check(_lib.lib().gcm_pe25_energy(dg.handle, ctypes.byref(s), _host.ptr(area), _host.ptr(out), _lib.stream()),
               "gcm_pe25_energy")
    ke, ate, geo = (float(x) for x in out.cpu())
    return ke, ate, geo, ke + ate + geo


def _minmax(t):
    out = _host.empty((3,))
    _lib.check(_lib.lib().gcm_diag_minmax(_host.ptr(t), t.numel(), _host.ptr(out), _lib.stream()), "gcm_diag_minmax")
    mn, mx, bad = (float(x) for x in out.cpu())
    return mn, mx, int(bad)


def solar_timestep(t, p, g, dt, utc, geom):
    """no_limits_2_5d.py:66-75: grey-radiation heating of the air columns and the ground over dt (t_lw = 0.1, t_sw =
    0.9, albedo = 0.3) -> (theta_n, GroundVars_n).  One launch (grey_solar.solar_timestep, csrc/physics.cu)."""
    from . import grey_solar
    t_n, gt_n = grey_solar.solar_timestep(t, p, g, dt, utc, geom)
    return t_n, GroundVars(gt_n, g.gw, g.snow, g.ice)


def full_timestep(p, u, v, t, q, g, dt, utc, geom, physics=False):
    """no_limits_2_5d.py:79-94: dynamics step + the STATS diagnostics (the reference's print is dropped).
    physics=True also runs the column physics the reference keeps below its early `return` (:96-103): solar_timestep on
    the new state."""
    p, u, v, t, q = matsuno_timestep(p, u, v, t, q, dt, geom)
    if physics:
        t, g = solar_timestep(t, p, g, dt, utc, geom)
    umin, umax, _ = _minmax(_host.dev(u))
    vmin, vmax, _ = _minmax(_host.dev(v))
    STATS["u_max"].append(umax)
    STATS["u_min"].append(umin)
    STATS["v_max"].append(vmax)
    STATS["v_min"].append(vmin)
    STATS["ke"].append(calc_energy(p, u, v, t, q, g, geom))
    return p, u, v, t, q, g


def run_model(height, width, layers, dt, timesteps, callback, stats=True, options=None):
    """no_limits_2_5d.py:220-236.  The state stays resident on the device for the whole run; it is
    downloaded only for `callback` and at the end.  stats=False skips the per-step diagnostics; options = kwargs of
    `dynamics.configure` (Coriolis, viscosity, flux-limited tracers; default: the reference's step)."""
    geom = gen_geometry(height, width, layers, sig_func=geometry.manabe_sig)
    if options:
        from .dynamics import configure
        configure(geom, **options)
    p, u, v, t, q, g = gen_initial_conditions(geom)
    v[0, 0, 0] = 0.1
    u *= 0
    st = Stepper(geom, p, u, v, t, q)
    for _ in range(timesteps):
        st.step(dt, 1)
        if stats:
            dp, du, dv, dtt, dq = st.tensors()
            umin, umax, _ = _minmax(du)
            vmin, vmax, _ = _minmax(dv)
            STATS["u_max"].append(umax)
            STATS["u_min"].append(umin)
            STATS["v_max"].append(vmax)
            STATS["v_min"].append(vmin)
            STATS["ke"].append(calc_energy(dp, du, dv, dtt, dq, g, geom))
        if callback:
            callback(*st.download())
    p, u, v, t, q = st.download()
    return p, u, v, t, q, g, geom


def save_checkpoint(path, p, u, v, t, q, g=None, utc=0.0, nsteps=0):
    """Checkpoint of a run (the reference has none: a run_model that dies starts over): the five
    prognostics as float64 SI magnitudes, the ground variables, the model time, the step count and the STATS
    diagnostics collected so far, in one .npz.  `p ... q` may be host arrays, Quantities or device tensors (e.g.
    `Stepper.tensors()`); restoring and stepping on gives bit for bit the run that was never interrupted."""
    def host(x):
        if isinstance(x, torch.Tensor):
            return x.detach().cpu().numpy()
        return np.asarray(_host.magnitude(x), dtype=np.float64)
    arrs = {k: host(x) for k, x in zip("puvtq", (p, u, v, t, q))}
    if g is not None:
        for k, x in zip(GroundVars._fields, g):
            arrs["g_" + k] = host(x)
    for k, vals in STATS.items():
        arrs["stats_" + k] = np.asarray(vals, dtype=np.float64)
    arrs["utc"] = np.float64(_host.scalar(utc))
    arrs["nsteps"] = np.int64(nsteps)
    with open(path, "wb") as f:
        np.savez(f, **arrs)


def load_checkpoint(path, restore_stats=True):
    """-> (p, u, v, t, q, g, utc, nsteps) as written by save_checkpoint (g is None if none was saved); STATS is
    restored to what it was at the checkpoint unless restore_stats=False."""
    with np.load(path) as z:
        state = tuple(z[k] for k in "puvtq")
        g = GroundVars(*(z["g_" + k] for k in GroundVars._fields)) if "g_gt" in z.files else None
        if restore_stats:
            STATS.clear()
            for k in z.files:
                if k.startswith("stats_"):
                    a = z[k]
                    STATS[k[6:]] = [tuple(r) for r in a] if a.ndim == 2 else [float(x) for x in a]
        return state + (g, float(z["utc"]), int(z["nsteps"]))


def main():
    """no_limits_2_5d.py:256-270 at a stable time step (the reference's 1800 s diverges)."""
    p, u, v, t, q, g, geom = run_model(8, 8, 3, 450.0, 200, None)
    print("pressures:", p * geom.sig + geom.ptop)


if __name__ == "__main__":
    main()
