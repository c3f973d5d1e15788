"""2.5-D sigma-layer primitive-equation Matsuno step, mirror of the reference `dynamics` module
(dynamics.py:15-237): same function names, positional signatures, [k, j, i] float64 layout and
out-of-place semantics; the arithmetic runs in hand-written sm_100a kernels behind the C ABI
(include/gcm_b200.h).  Inputs may be pint Quantities, numpy arrays (SI) or torch tensors on the device;
results come back in the same family.

Resident use (no host round trip per step): `Stepper` keeps the state, the geometry tables and the
scratch fields on the device and advances `nsteps` Matsuno steps per call.
"""
import ctypes

import torch

from . import _abi, _host, _lib
from .geometry import _push_options, device_geom

__all__ = ["calc_pu", "calc_pv", "un_pu", "un_pv", "aflux", "advec_sig", "advec_m_pu", "compute_geopotential",
           "pgf", "advec_t", "half_timestep", "matsuno_timestep", "Stepper", "StepOptions", "configure"]

_UNITS = ("pascal", "meter / second", "meter / second", "kelvin", "dimensionless")


class StepOptions:
    """Opt-in terms of the 2.5-D half step (include/gcm_b200.h gcm_pe25_options).
    All off = the reference's step.  coriolis: dynamics.py:82-95 switched on; viscosity: kinematic horizontal
    viscosity in m2/s (viscosity.py:12-25 on the lat-lon metric); limit_q / limit_t: van Leer flux-limited horizontal
    advection of q / theta (flux_limiter.py:10-32 composed into advec_t, the TODO at dynamics.py:217-218)."""

    def __init__(self, coriolis=False, viscosity=0.0, limit_q=False, limit_t=False):
        self.coriolis, self.limit_q, self.limit_t = bool(coriolis), bool(limit_q), bool(limit_t)
        self.viscosity = _host.scalar(viscosity)
        if not self.viscosity >= 0.0:
            raise ValueError("viscosity must be >= 0 (m2/s)")

    def any(self):
        return self.coriolis or self.limit_q or self.limit_t or self.viscosity != 0.0


def configure(geom, coriolis=False, viscosity=0.0, limit_q=False, limit_t=False):
    """Switch the opt-in terms of `half_timestep` / `matsuno_timestep` / `Stepper.step` on or off for `geom`
    (whole grids, and latitude bands created afterwards by `bands.BandStepper`, which then carry two halo rows on
    either side).  `configure(geom)` restores the reference's step.  Returns the StepOptions."""
    opt = StepOptions(coriolis, viscosity, limit_q, limit_t)
    resident = list(geom._dev.values())
    # validate every resident geometry BEFORE touching anything: the opt-in terms need exactly 2 + 2 halo rows on a
    # band (the one-exchange bands carry 2 + 4 and run the row-segment schedule, which the options do not support)
    for dg in resident:
        if not dg.wrap_j and opt.any() != dg.options_on and not (dg.row_lo == 2 and dg.H - dg.row_hi == 2):
            raise ValueError("a latitude band of this geometry is already resident with a halo layout other than 2 + 2 "
                             "rows: configure() before creating the BandStepper")
    geom.step_options = opt
    for dg in resident:
        if dg.wrap_j or (dg.row_lo == 2 and dg.H - dg.row_hi == 2):
            _push_options(geom, dg)
    return opt


def _struct(ts):
    return _abi.State(*[_host.ptr(t) for t in ts])


def _workspace(dg, nbatch, owner=None):
    """Scratch fields of a half step.  A Stepper / BandStepper owns its own (`owner`), so two steppers of one geometry
    driven from different streams do not share spu / pgf / the star state; the functional calls (half_timestep,
    matsuno_timestep, ...) share one per device geometry and are single-stream per geometry by contract.  `dg._ws`
    always names the workspace of the latest call (the tests read work fields through it)."""
    holder = owner if owner is not None else dg
    ws = getattr(holder, "_ws_own", None)
    need = _lib.lib().gcm_pe25_workspace_bytes(dg.handle, nbatch)
    if ws is None or ws.numel() * 8 < need or ws.device != _lib.device():
        ws = torch.empty((need + 7) // 8, dtype=torch.float64, device=_lib.device())
        holder._ws_own = ws
    dg._ws = ws
    return ws, need


def _shape_check(dg, p, u, v, t, q):
    L, H, W = dg.L, dg.H, dg.W
    if tuple(p.shape[-2:]) != (H, W):
        raise ValueError("p has shape %s, geometry is %d x %d" % (tuple(p.shape), H, W))
    for name, a in zip("uvtq", (u, v, t, q)):
        if tuple(a.shape[-3:]) != (L, H, W):
            raise ValueError("%s has shape %s, geometry is %d x %d x %d" % (name, tuple(a.shape), L, H, W))
    nb = 1 if p.dim() == 2 else int(p.shape[0])
    return nb


class Stepper:
    """Device-resident 2.5-D model state advanced by `gcm_pe25_matsuno_step` (dynamics.py:230-237).

    >>> st = Stepper(geom, p, u, v, t, q)        # one upload
    >>> st.step(dt, nsteps=100)                  # 100 Matsuno steps, state never leaves HBM
    >>> p, u, v, t, q = st.download()
    A leading ensemble dimension ([b, j, i] / [b, k, j, i]) steps `b` independent members at once.
    """

    def __init__(self, geom, p, u, v, t, q):
        self.geom = geom
        self.dg = device_geom(geom)
        self.family = _host.Family(p, u, v, t, q)
        self.cur = [_host.dev(x).clone() for x in (p, u, v, t, q)]
        self.nbatch = _shape_check(self.dg, *self.cur)
        self.nxt = [torch.empty_like(x) for x in self.cur]
        self.nsteps_done = 0

    def upload(self, p, u, v, t, q):
        for dst, src in zip(self.cur, (p, u, v, t, q)):
            dst.copy_(_host.dev(src), non_blocking=True)

    def step(self, dt, nsteps=1):
        if getattr(self, "_pipe", None) is not None:
            self.host_join()        # pending copy-outs of pipelined host steps read the buffers this step writes
        ws, need = _workspace(self.dg, self.nbatch, self)
        sin, sout = _struct(self.cur), _struct(self.nxt)
        _lib.check(_lib.lib().gcm_pe25_matsuno_step(self.dg.handle, ctypes.byref(sin), ctypes.byref(sout),
                                                    _host.scalar(dt), int(nsteps), self.nbatch, _host.ptr(ws), need,
                                                    _lib.stream()), "gcm_pe25_matsuno_step")
        self.cur, self.nxt = self.nxt, self.cur
        self.nsteps_done += int(nsteps)

    def step_host(self, host_in, host_out, dt, nsteps=1, pipelined=False):
        """Host-resident caller: copy the five (pinned) host tensors `host_in` to the device, advance, copy the
        new state into the five (pinned) host tensors `host_out`.  All asynchronous on the current stream.
        One step of one member goes through `gcm_pe25_matsuno_step_host`: latitude blocks copied in, stepped and
        copied out on three streams, so the PCIe link runs in both directions at once.

        pipelined=True (a time loop whose state lives in host memory between steps): consecutive calls overlap -- the
        copy-in of the next step starts while the last blocks of this one are still on their way out
        (`gcm_pe25_matsuno_step_host_pipelined`).  `host_out` is then complete only after `host_join()`."""
        ok_host = all(not x.is_cuda and x.is_contiguous() and x.dtype == torch.float64 for x in list(host_in) + list(host_out))
        if pipelined and int(nsteps) == 1 and self.nbatch == 1 and self.cur[0].dim() == 2 and not self.dg.options_on and ok_host:
            if getattr(self, "_pipe", None) is None:
                mk = lambda: [torch.empty_like(x) for x in self.cur]
                self._pipe = {"x": (mk(), mk()), "y": (self.cur, self.nxt), "star": mk(), "n": 0}
            pp = self._pipe
            n = pp["n"]
            ws, need = _workspace(self.dg, 1, self)
            hi, ho = _struct(host_in), _struct(host_out)
            sx, ss, sy = _struct(pp["x"][n % 2]), _struct(pp["star"]), _struct(pp["y"][n % 2])
            st = _lib.lib().gcm_pe25_matsuno_step_host_pipelined(self.dg.handle, ctypes.byref(hi), ctypes.byref(ho),
                                                                 ctypes.byref(sx), ctypes.byref(ss), ctypes.byref(sy),
                                                                 _host.scalar(dt), 0, n, _host.ptr(ws), need, _lib.stream())
            if st == 0:
                pp["n"] = n + 1
                self.cur = pp["y"][n % 2]          # the newest state on the device
                self.nxt = pp["y"][(n + 1) % 2]
                self.nsteps_done += 1
                return
            if st != -4:                           # GCM_EUNSUP: no row-segment kernels for this geometry -> plain path
                _lib.check(st, "gcm_pe25_matsuno_step_host_pipelined")
        if int(nsteps) == 1 and self.nbatch == 1 and self.cur[0].dim() == 2 and not self.dg.options_on and all(
                not x.is_cuda and x.is_contiguous() and x.dtype == torch.float64 for x in list(host_in) + list(host_out)):
            if getattr(self, "_star", None) is None:
                self._star = [torch.empty_like(x) for x in self.cur]
            ws, need = _workspace(self.dg, 1, self)
            hi, ho = _struct(host_in), _struct(host_out)
            sc, ss, sn = _struct(self.cur), _struct(self._star), _struct(self.nxt)
            _lib.check(_lib.lib().gcm_pe25_matsuno_step_host(self.dg.handle, ctypes.byref(hi), ctypes.byref(ho),
                                                             ctypes.byref(sc), ctypes.byref(ss), ctypes.byref(sn),
                                                             _host.scalar(dt), 0, _host.ptr(ws), need, _lib.stream()),
                       "gcm_pe25_matsuno_step_host")
            self.cur, self.nxt = self.nxt, self.cur
            self.nsteps_done += 1
            return
        for dst, src in zip(self.cur, host_in):
            dst.copy_(src, non_blocking=True)
        self.step(dt, nsteps)
        for dst, src in zip(host_out, self.cur):
            dst.copy_(src, non_blocking=True)

    def host_join(self):
        """Make the current stream wait for the copy-outs of every `step_host(..., pipelined=True)` issued so far."""
        _lib.check(_lib.lib().gcm_host_pipe_join(_lib.stream()), "gcm_host_pipe_join")

    def tensors(self):
        """The current state as device tensors (p, u, v, t, q); valid until the next step()."""
        return tuple(self.cur)

    def download(self):
        return tuple(self.family.out(x, unit) for x, unit in zip(self.cur, _UNITS))


def half_timestep(p, u, v, t, q, sp, su, sv, st, sq, dt, geom):
    """dynamics.py:183-227: out = base + dt * F(star) for all five prognostics."""
    fam = _host.Family(p, u, v, t, q, sp, su, sv, st, sq)
    dg = device_geom(geom)
    base = [_host.dev(x) for x in (p, u, v, t, q)]
    star = [_host.dev(x) for x in (sp, su, sv, st, sq)]
    nb = _shape_check(dg, *base)
    _shape_check(dg, *star)
    out = [torch.empty_like(x) for x in base]
    ws, need = _workspace(dg, nb)
    sb, ss, so = _struct(base), _struct(star), _struct(out)
    _lib.check(_lib.lib().gcm_pe25_half_step(dg.handle, ctypes.byref(sb), ctypes.byref(ss), ctypes.byref(so),
                                             _host.scalar(dt), nb, _host.ptr(ws), need, _lib.stream()),
               "gcm_pe25_half_step")
    return tuple(fam.out(x, unit) for x, unit in zip(out, _UNITS))


def matsuno_timestep(p, u, v, t, q, dt, geom, boundary_conditions=None):
    """dynamics.py:230-237: predictor X* = X + dt F(X), corrector X' = X + dt F(X*)."""
    if boundary_conditions:
        sp, su, sv, st, sq = half_timestep(p, u, v, t, q, p, u, v, t, q, dt, geom)
        sp, su, sv, st, sq = boundary_conditions(sp, su, sv, st, sq, dt, geom)
        op, ou, ov, ot, oq = half_timestep(p, u, v, t, q, sp, su, sv, st, sq, dt, geom)
        return boundary_conditions(op, ou, ov, ot, oq, dt, geom)
    fam = _host.Family(p, u, v, t, q)
    dg = device_geom(geom)
    base = [_host.dev(x) for x in (p, u, v, t, q)]
    nb = _shape_check(dg, *base)
    out = [torch.empty_like(x) for x in base]
    ws, need = _workspace(dg, nb)
    sb, so = _struct(base), _struct(out)
    _lib.check(_lib.lib().gcm_pe25_matsuno_step(dg.handle, ctypes.byref(sb), ctypes.byref(so), _host.scalar(dt), 1, nb,
                                                _host.ptr(ws), need, _lib.stream()), "gcm_pe25_matsuno_step")
    return tuple(fam.out(x, unit) for x, unit in zip(out, _UNITS))


# ---- the operators half_timestep is built from (each is its own kernel behind the C ABI) ---------------
def _op3(name, geom, ins, nout=1, shapes=None, extra=()):
    fam = _host.Family(*ins)
    dg = device_geom(geom)
    ts = [_host.dev(x) for x in ins]
    L, H, W = dg.L, dg.H, dg.W
    outs = [_host.empty(s) for s in (shapes or [(L, H, W)] * nout)]
    args = [dg.handle] + [_host.ptr(x) for x in ts] + [_host.ptr(x) for x in outs] + list(extra) + [_lib.stream()]
    _lib.check(getattr(_lib.lib(), name)(*args), name)
    res = tuple(fam.out(x) for x in outs)
    return res[0] if len(res) == 1 else res


def calc_pu(p, u):
    """pu = u * iph(p)   (dynamics.py:15-17).  No geometry in the reference signature: periodic in i."""
    return _flux(0, p, u)


def calc_pv(p, v):
    """pv = v * jph(p)   (dynamics.py:20-22)."""
    return _flux(1, p, v)


def un_pu(pu, p):
    """u = pu / iph(p)   (dynamics.py:25-27)."""
    return _flux(2, p, pu)


def un_pv(pv, p):
    """v = pv / jph(p)   (dynamics.py:30-32)."""
    return _flux(3, p, pv)


class _Grid:
    """Just the extents: calc_pu & co take no geometry in the reference."""

    def __init__(self, L, H, W):
        import numpy as np
        self.layers, self.height, self.width = L, H, W
        self.sige = np.linspace(1, 0, L + 1)
        self.sigb, self.sigt = self.sige[:-1], self.sige[1:]
        self.dsig = self.sigb - self.sigt
        self.sig = (self.sigb + self.sigt) / 2
        self.dx_j = self.dx_h = np.ones(H)
        self.dy, self.ptop = 1.0, 0.0
        self.heightmap = np.zeros((H, W))
        self._dev = {}


_GRIDS = {}


def _flux(op, p, a):
    fam = _host.Family(p, a)
    tp, ta = _host.dev(p), _host.dev(a)
    if ta.dim() == 2:
        ta = ta.unsqueeze(0)
    L, H, W = ta.shape
    key = (L, H, W, str(_lib.device()))
    if key not in _GRIDS:
        if len(_GRIDS) > 16:
            _GRIDS.clear()
        _GRIDS[key] = _Grid(L, H, W)
    g = _GRIDS[key]
    if W % 2 and W != 1:
        raise ValueError("odd W is not supported by the device geometry (low_pass.py:57)")
    dg = device_geom(g)
    out = _host.empty((L, H, W))
    fn = ("gcm_pe25_calc_pu", "gcm_pe25_calc_pv", "gcm_pe25_un_pu", "gcm_pe25_un_pv")[op]
    first, second = (tp, ta) if op < 2 else (ta, tp)
    _lib.check(getattr(_lib.lib(), fn)(dg.handle, _host.ptr(first), _host.ptr(second), _host.ptr(out), _lib.stream()), fn)
    return fam.out(out)


def aflux(pu, pv, geom):
    """dynamics.py:35-46 -> (pit [H, W], sd [L, H, W])."""
    return _op3("gcm_pe25_aflux", geom, (pu, pv), shapes=[(geom.height, geom.width), (geom.layers, geom.height, geom.width)])


def advec_sig(sd, q, geom):
    """dynamics.py:49-52."""
    return _op3("gcm_pe25_advec_sig", geom, (sd, q))


def advec_m_pu(p, u, v, pu, pv, geom):
    """dynamics.py:55-108 -> (dut, dvt)."""
    return _op3("gcm_pe25_advec_m_pu", geom, (p, u, v, pu, pv), nout=2)


def compute_geopotential(p, t, geom):
    """dynamics.py:111-142 (the reference's two prints per call are dropped)."""
    return _op3("gcm_pe25_geopotential", geom, (p, t))


def pgf(p, t, geom):
    """dynamics.py:147-171 -> (pgfu, pgfv, phiu, phiv)."""
    n3 = geom.layers * geom.height * geom.width
    ws = _host.empty((2 * n3,))
    return _op3("gcm_pe25_pgf", geom, (p, t), nout=4, extra=(_host.ptr(ws), 2 * n3 * 8))


def advec_t(pu, pv, t, geom):
    """dynamics.py:174-181."""
    return _op3("gcm_pe25_advec_t", geom, (pu, pv, t))
