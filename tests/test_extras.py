"""Opt-in terms of the 2.5-D half step: Coriolis, horizontal viscosity, van Leer
flux-limited tracer advection.  Default OFF; the default step stays the reference's.

Pinning: the Coriolis term is pinned bit for bit against the reference run with its dead branch (dynamics.py:82)
switched on in memory (oracle/make_golden_ext.py -> tests/golden/run25_24x36x9_coriolis.npz).  The limiter and the
viscosity have no reference counterpart inside the 2.5-D step (TODOs at dynamics.py:217-218): their oracle
restatement (np_oracle.half_timestep_ext) is "parity unpinned" and is checked here through properties -- reduction to
the reference's own operators, conservation, monotonicity -- and the kernels are then compared with that oracle.

Tolerances: the kernels add the terms as a correction to the state the half step has written (one extra launch), the
oracle adds them inside the tendency sums, so the results agree to round-off, not bit for bit:
single call 1e-12 * max|field|, N-step runs 1e-11 (u, v against max(|u|, |v|)).
"""
import ctypes

import numpy as np
import pytest

import np_oracle as O
from conftest import load_golden
from gcmiipy_b200 import _lib, dynamics, geometry
from test_parity import check_state, rel

TOL_CALL = 1e-12
TOL_RUN = 1e-11


# ---- the oracle itself (no kernel involved) -------------------------------------------------------------------
def test_oracle_coriolis_is_pinned_to_the_reference():
    g = load_golden("run25_24x36x9_coriolis")
    og = O.gen_geometry(24, 36, 9, sig_func=O.manabe_sig)
    s = tuple(g[k + "_0"] for k in "puvtq")
    opt = O.StepOptions(coriolis=True)
    cu, cv = O.coriolis_terms(g["pu"], g["pv"], og)
    assert np.array_equal(cu, g["cor_u"]) or rel(cu, g["cor_u"]) < 1e-15   # cor_* are stored as differences of sums
    assert rel(cv, g["cor_v"]) < 1e-15
    dut, dvt = O.advec_m_pu(s[0], s[1], s[2], g["pu"], g["pv"], og)
    assert np.array_equal(dut + cu, g["dut"]) and np.array_equal(dvt + cv, g["dvt"])
    hs = O.half_timestep_ext(*s, *s, float(g["dt"]), og, opt)
    for a, k in zip(hs, "puvtq"):
        assert np.array_equal(a, g["hs_" + k]), k
    cur = s
    for i in range(1, 11):
        cur = O.matsuno_timestep_ext(*cur, float(g["dt"]), og, opt)
        if i in (1, 10):
            for a, k in zip(cur, "puvtq"):
                assert np.array_equal(a, g["%s_%d" % (k, i)]), (k, i)


def test_oracle_options_off_is_the_reference_step():
    og = O.gen_geometry(24, 36, 9, sig_func=O.manabe_sig)
    s = O.synthetic_state(og)
    for a, b in zip(O.half_timestep(*s, *s, 450.0, og), O.half_timestep_ext(*s, *s, 450.0, og, O.StepOptions())):
        assert np.array_equal(a, b)


def test_oracle_laplacian_reduces_to_viscosity_py():
    """dx == dy: the metric Laplacian is viscosity.finite_laplacian_2d (viscosity.py:12-19) up to round-off."""
    q = np.random.default_rng(3).standard_normal((4, 12, 10))
    a, b = O.laplacian_h(q, 250e3, 250e3), O.finite_laplacian_2d(q, 250e3)
    assert rel(a, b) < 1e-14


def test_oracle_limited_edge_reductions():
    rng = np.random.default_rng(5)
    q = np.cumsum(np.ones((3, 6, 16)), -1) * 2.5 + 7           # linear in i: r = 1, phi = 1 -> centred value
    f = rng.standard_normal(q.shape)
    e = O.limited_edge_value(q, f, -1)
    assert np.allclose(e[..., 2:-2], O.iph(q)[..., 2:-2], rtol=1e-15)
    q = rng.standard_normal((3, 6, 16))                         # bounded by the two cells of the edge (phi in [0, 2])
    for ax in (-1, -2):
        e = O.limited_edge_value(q, f, ax)
        lo, hi = np.minimum(q, np.roll(q, -1, ax)), np.maximum(q, np.roll(q, -1, ax))
        assert np.all(e >= lo - 1e-15) and np.all(e <= hi + 1e-15)
    # 1-D building blocks of flux_limiter.py: same slope ratio / limiter as calc_r / van_leer
    row = rng.standard_normal(16)
    b = np.roll(row, -1) - row
    e1 = O.limited_edge_value(row, np.ones(16), 0)
    assert np.array_equal(e1, row + 0.5 * O.van_leer(O.calc_r(row)) * b)
    assert np.array_equal(O.limited_edge_value(row, -np.ones(16), 0)[b == 0], np.roll(row, -1)[b == 0])


def test_oracle_limited_advection_is_monotone_and_conservative():
    """The square pulse of flux_limiter.py:74-79 in uniform flow: the limited flux creates no new extrema and
    conserves the total; the centred flux of advec_t (dynamics.py:174-181) overshoots."""
    n, dx, dt, u = 64, 1.0, 0.2, 1.0
    q0 = np.zeros(n); q0[20:30] = 1.0
    flux = np.full(n, u)

    def run(edge):
        q = q0.copy()
        for _ in range(100):
            # Matsuno predictor-corrector like the model (dynamics.py:230-237)
            f = flux * edge(q)
            qs = q + (np.roll(f, 1) - f) * dt / dx
            f = flux * edge(qs)
            q = q + (np.roll(f, 1) - f) * dt / dx
        return q

    lim = run(lambda q: O.limited_edge_value(q, flux, 0))
    cen = run(lambda q: (q + np.roll(q, -1)) / 2)
    assert abs(lim.sum() - q0.sum()) < 1e-12 and abs(cen.sum() - q0.sum()) < 1e-12
    assert lim.min() > -1e-3 and lim.max() < 1 + 1e-3, (lim.min(), lim.max())
    assert cen.min() < -0.05 or cen.max() > 1.05                 # what the limiter is for
    tv = lambda q: np.abs(np.roll(q, -1) - q).sum()
    assert tv(lim) <= tv(q0) + 1e-3


def test_oracle_limiter_conserves_what_the_centred_flux_conserves():
    og = O.gen_geometry(24, 36, 9, sig_func=O.manabe_sig)
    s = O.synthetic_state(og)
    rng = np.random.default_rng(11)
    s = s[:4] + (s[4] * (1 + 0.3 * rng.random(s[4].shape)),)     # rough tracer: the limiter has work to do
    a = O.half_timestep(*s, *s, 450.0, og)
    b = O.half_timestep_ext(*s, *s, 450.0, og, O.StepOptions(limit_q=True, limit_t=True))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for f in (3, 4):                                             # layer totals of pi * tracer
        ta, tb = (a[f] * a[0]).sum((-1, -2)), (b[f] * b[0]).sum((-1, -2))
        assert np.max(np.abs(ta - tb) / np.abs(ta)) < 1e-13
        assert np.max(np.abs(a[f] - b[f])) > 0


# ---- kernels against the oracle ----------------------------------------------------------------------------------
OPTS = {
    "coriolis": dict(coriolis=True),
    "viscosity": dict(viscosity=2.0e5),
    "limit_q": dict(limit_q=True),
    "limit_t": dict(limit_t=True),
    "all": dict(coriolis=True, viscosity=1.0e5, limit_q=True, limit_t=True),
}


def _oopt(kw):
    return O.StepOptions(kw.get("coriolis", False), kw.get("viscosity", 0.0), kw.get("limit_q", False),
                         kw.get("limit_t", False))


def _case(H, W, L, seed=7, rough=True):
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    s = O.synthetic_state(og, seed=seed)
    if rough:
        rng = np.random.default_rng(seed)
        s = s[:4] + (s[4] * (1 + 0.3 * rng.random(s[4].shape)),)
    return geom, og, s


def test_coriolis_golden_through_the_kernels(backend):
    """Reference with dynamics.py:82 switched on vs the kernels: half step and 10 Matsuno steps."""
    g = load_golden("run25_24x36x9_coriolis")
    geom = geometry.gen_geometry(24, 36, 9, sig_func=geometry.manabe_sig)
    dynamics.configure(geom, coriolis=True)
    s = tuple(g[k + "_0"] for k in "puvtq")
    dt = float(g["dt"])
    check_state(dynamics.half_timestep(*s, *s, dt, geom), tuple(g["hs_" + k] for k in "puvtq"), TOL_CALL)
    st = dynamics.Stepper(geom, *s)
    st.step(dt, 1)
    check_state(st.download(), tuple(g[k + "_1"] for k in "puvtq"), TOL_CALL)
    st.step(dt, 9)
    check_state(st.download(), tuple(g[k + "_10"] for k in "puvtq"), TOL_RUN)
    dynamics.configure(geom)                                    # off again: the reference's step, same geometry object
    base = load_golden("ops25_24x36x9")
    s0 = tuple(base[k] for k in "puvtq")
    check_state(dynamics.half_timestep(*s0, *s0, 450.0, geom), tuple(base["hs_" + k] for k in "puvtq"), TOL_CALL)


@pytest.mark.parametrize("name", sorted(OPTS))
@pytest.mark.parametrize("H,W,L", [(24, 36, 9), (10, 64, 9), (9, 14, 4)])
def test_half_step_with_options_vs_oracle(backend, name, H, W, L):
    """36: narrow fused kernels; 64: tiled update; 14 x 4 layers: the general 4-kernel path (W has a factor 7)."""
    geom, og, s = _case(H, W, L)
    dynamics.configure(geom, **OPTS[name])
    star = O.half_timestep(*s, *s, 300.0, og)                    # a star state that differs from the base
    got = dynamics.half_timestep(*s, *star, 300.0, geom)
    ref = O.half_timestep_ext(*s, *star, 300.0, og, _oopt(OPTS[name]))
    check_state(got, ref, TOL_CALL)
    plain = O.half_timestep(*s, *star, 300.0, og)
    assert any(np.max(np.abs(a - b)) > 0 for a, b in zip(ref, plain))     # the option does something


@pytest.mark.parametrize("H,W,L,dt,n", [(24, 36, 9, 450.0, 8), (46, 72, 9, 225.0, 6), (12, 96, 9, 200.0, 3)])
def test_runs_with_all_options_vs_oracle(backend, H, W, L, dt, n):
    geom, og, s = _case(H, W, L, seed=H + W)
    kw = OPTS["all"]
    dynamics.configure(geom, **kw)
    ref = s
    for _ in range(n):
        ref = O.matsuno_timestep_ext(*ref, dt, og, _oopt(kw))
    assert all(np.isfinite(a).all() for a in ref)
    st = dynamics.Stepper(geom, *s)
    st.step(dt, n)                                               # n >= 5 on small grids replays the step-pair graph
    check_state(st.download(), ref, TOL_RUN)
    one = dynamics.matsuno_timestep(*s, dt, geom)
    check_state(one, O.matsuno_timestep_ext(*s, dt, og, _oopt(kw)), TOL_CALL)


def test_options_on_an_ensemble(backend):
    geom, og, s = _case(24, 36, 9)
    kw = OPTS["all"]
    dynamics.configure(geom, **kw)
    _, _, s2 = _case(24, 36, 9, seed=99)
    batch = tuple(np.stack([a, b]) for a, b in zip(s, s2))
    st = dynamics.Stepper(geom, *batch)
    st.step(450.0, 2)
    got = st.download()
    for m, sm in enumerate((s, s2)):
        ref = sm
        for _ in range(2):
            ref = O.matsuno_timestep_ext(*ref, 450.0, og, _oopt(kw))
        check_state(tuple(a[m] for a in got), ref, TOL_RUN)


def test_options_are_refused_where_they_do_not_apply(backend):
    geom, og, s = _case(24, 36, 9)
    dynamics.configure(geom, limit_q=True)
    with pytest.raises(ValueError):                      # a band with the reference's halo widths (1 north, 2 south)
        geometry.device_geom(geom, band=(0, 12, 1, 2))
    dg = geometry.device_geom(geom)
    assert dg.options_on
    # row-segment entry point (bands, host-resident pipeline): GCM_EUNSUP while an option is on
    st = dynamics.Stepper(geom, *s)
    ws, need = dynamics._workspace(dg, 1)
    seg = (ctypes.c_int * 4)(0, 24, 0, 0)
    sc, sn = dynamics._struct(st.cur), dynamics._struct(st.nxt)
    from gcmiipy_b200 import _host
    rc = _lib.lib().gcm_pe25_half_step_rows(dg.handle, ctypes.byref(sc), ctypes.byref(sc), ctypes.byref(sn), 450.0, 1,
                                            _host.ptr(ws), need, seg, seg, _lib.stream())
    assert rc == -4
    # a negative viscosity is an argument error
    with pytest.raises(ValueError):
        dynamics.configure(geom, viscosity=-1.0)
    # the host-resident step falls back to copy-in / step / copy-out and still applies the option
    import torch
    hin = [torch.from_numpy(np.ascontiguousarray(a)) for a in s]
    hout = [torch.empty_like(a) for a in hin]
    st.step_host(hin, hout, 450.0)
    if backend == "gpu":
        torch.cuda.synchronize()
    ref = O.matsuno_timestep_ext(*s, 450.0, og, O.StepOptions(limit_q=True))
    check_state(tuple(a.numpy() for a in hout), ref, TOL_CALL)


def test_pressure_from_heightmap_golden():
    """geometry.pressure_from_heightmap (geometry.py:185-231), host-side initial-condition helper."""
    g = load_golden("barometric")
    assert np.array_equal(O.pressure_from_heightmap(g["height"], float(g["p0"]), float(g["t0"])), g["p"])
    assert rel(geometry.pressure_from_heightmap(g["height"], float(g["p0"]), float(g["t0"])), g["p"]) < 1e-15


def test_checkpoint_resume_is_bit_identical(backend, tmp_path):
    """6 steps straight == 3 steps, checkpoint, fresh Stepper from the file, 3 steps."""
    from gcmiipy_b200 import no_limits_2_5d as nl
    geom, og, s = _case(24, 36, 9, seed=21)
    dynamics.configure(geom, coriolis=True, limit_q=True)
    a = dynamics.Stepper(geom, *s)
    a.step(450.0, 6)
    b = dynamics.Stepper(geom, *s)
    b.step(450.0, 3)
    nl.STATS.clear()
    nl.STATS["u_max"].append(1.5)
    nl.STATS["ke"].append((1.0, 2.0, 3.0, 6.0))
    ground = nl.gen_initial_conditions(geom)[5]
    path = str(tmp_path / "ck.npz")
    nl.save_checkpoint(path, *b.tensors(), g=ground, utc=3 * 450.0, nsteps=3)
    nl.STATS.clear()
    p, u, v, t, q, g2, utc, n = nl.load_checkpoint(path)
    assert (utc, n) == (1350.0, 3) and nl.STATS["u_max"] == [1.5] and nl.STATS["ke"] == [(1.0, 2.0, 3.0, 6.0)]
    assert np.array_equal(g2.gt, ground.gt)
    c = dynamics.Stepper(geom, p, u, v, t, q)
    c.step(450.0, 3)
    for x, y in zip(a.download(), c.download()):
        assert np.array_equal(x, y)
    nl.STATS.clear()


@pytest.mark.gpu
def test_options_full_size_properties():
    """1 x 1.25 deg grid (BASELINE configs[2]): with the limiter on q only, surface pressure, winds and theta are
    those of the default run bit for bit and the global total of pi * q * dsig is the default's (flux form: the
    limiter moves tracer between cells, never creates it); with every option on the state stays finite.
    nu = 1e3 m2/s: the explicit diffusion limit nu dt / dx_j^2 < 1/4 at the row next to the pole (dx_j = 1.2 km)."""
    import torch
    H, W, L, dt = 180, 288, 9, 60.0
    geom, og, s = _case(H, W, L, seed=3)
    st = dynamics.Stepper(geom, *s)
    st.step(dt, 5)
    base = st.download()
    dynamics.configure(geom, limit_q=True)
    st = dynamics.Stepper(geom, *s)
    st.step(dt, 5)
    lim = st.download()
    for f in (0, 1, 2, 3):
        assert np.array_equal(base[f], lim[f])
    dsig = og.dsig.reshape(-1)
    ta, tb = ((base[4] * base[0]).sum((-1, -2)) * dsig).sum(), ((lim[4] * lim[0]).sum((-1, -2)) * dsig).sum()
    assert abs(ta - tb) / abs(ta) < 1e-12
    assert np.max(np.abs(base[4] - lim[4])) > 0
    dynamics.configure(geom, coriolis=True, viscosity=1.0e3, limit_q=True, limit_t=True)
    st = dynamics.Stepper(geom, *s)
    st.step(dt, 20)
    torch.cuda.synchronize()
    assert all(np.isfinite(a).all() for a in st.download())


@pytest.mark.parametrize("H,W,L,ptop", [(1, 16, 17, 0.0), (2, 8, 3, 0.0), (3, 2, 3, 0.0), (5, 6, 9, 1000.0), (8, 8, 3, 0.0)])
def test_options_on_degenerate_grids_vs_oracle(backend, H, W, L, ptop):
    """One and two rows (every j neighbour is the row itself / the other row), two columns (i +- 2 wraps onto the
    cell), 17 layers (the general 4-kernel path), ptop != 0, a mountain: the periodic wraps of the limiter's
    five-point reach and of the Coriolis / viscosity stencils against the numpy rolls of the oracle."""
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    geom.ptop = og.ptop = ptop
    hm = np.zeros((H, W)); hm[0, W // 2] = 800.0
    geom.heightmap = hm; og.heightmap = hm.copy()
    s = O.synthetic_state(og, seed=H * 100 + W)
    rng = np.random.default_rng(H + W)
    s = s[:4] + (s[4] * (1 + 0.3 * rng.random(s[4].shape)),)
    kw = dict(coriolis=True, viscosity=5.0e4, limit_q=True, limit_t=True)
    dynamics.configure(geom, **kw)
    dt = 100.0
    ref = O.matsuno_timestep_ext(*s, dt, og, _oopt(kw))
    assert all(np.isfinite(a).all() for a in ref)
    check_state(dynamics.matsuno_timestep(*s, dt, geom), ref, TOL_CALL)


def test_coriolis_on_a_square_geometry_is_zero(backend):
    """gen_square_geometry has no latitudes (geometry.py:174: lat = 0): the Coriolis parameters are zero and the
    option changes nothing but round-off-free zeros."""
    geom = geometry.gen_square_geometry(6, 10, 3, 300e3, 250e3, sig_func=geometry.manabe_sig)
    og = O.gen_square_geometry(6, 10, 3, 300e3, 250e3, sig_func=O.manabe_sig)
    s = O.synthetic_state(og, seed=5)
    base = dynamics.matsuno_timestep(*s, 200.0, geom)
    dynamics.configure(geom, coriolis=True)
    cor = dynamics.matsuno_timestep(*s, 200.0, geom)
    for a, b in zip(base, cor):
        assert np.array_equal(a, b)
    check_state(cor, O.matsuno_timestep_ext(*s, 200.0, og, O.StepOptions(coriolis=True)), TOL_CALL)


def test_run_model_with_options(backend):
    """no_limits_2_5d.run_model (no_limits_2_5d.py:220-236) with the opt-in terms: same ICs, same loop."""
    from gcmiipy_b200 import no_limits_2_5d as nl
    kw = dict(coriolis=True, limit_q=True, viscosity=1.0e5)
    p, u, v, t, q, g, geom = nl.run_model(8, 8, 3, 450.0, 4, None, stats=False, options=kw)
    og = O.gen_geometry(8, 8, 3, sig_func=O.manabe_sig)
    ref = O.run_model_ic(og)
    for _ in range(4):
        ref = O.matsuno_timestep_ext(*ref, 450.0, og, _oopt(kw))
    check_state((p, u, v, t, q), ref, TOL_RUN)
