"""Parity of the kernel path with the reference: (1) the golden vectors produced by the UNMODIFIED reference
(tests/golden, oracle/make_golden.py) and (2) the pinned numpy oracle on seeded inputs.

Every case runs on two builds of the same kernel sources (conftest.backend): "gpu" = the sm_100a library on a
B200 through the C ABI (-m gpu, the parity tests proper) and "emu" = the sources compiled for the CPU emulator
(no GPU in the build container).

Tolerances (fp64, relative to max|ref| of the field; u and v against max(|u|, |v|) because u starts at 0):
  * + - * / only kernels (shallow water, Laplacian, flux limiter, shifts, aflux, advec_*): BIT-EXACT
  * kernels with pow() or the FFT filter, single call: 1e-13
  * phiu / phiv (differences of a geopotential of magnitude ~3e5 m2/s2): 1e-13 of p * max|phi| / dx
  * N-step runs at the stable time steps of SURVEY.md section 4: 1e-11
"""
import numpy as np
import pytest

import np_oracle as O
from conftest import load_golden
from gcmiipy_b200 import (coordinates, coordinates_1d, coordinates_3d, dynamics, flux_limiter, geometry, low_pass,
                          matsumo_temp, matsuno_c_grid, no_limits_2_5d, no_limits_2d, phi_port, temperature,
                          viscosity)

TOL_CALL = 1e-13
TOL_RUN = 1e-11


def rel(a, b, scale=None):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    s = np.max(np.abs(b)) if scale is None else scale
    return np.max(np.abs(a - b)) / max(s, 1e-300)


def exact(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b), "max |diff| = %g" % np.max(np.abs(a - b))


def check_state(got, ref, tol):
    suv = max(np.max(np.abs(ref[1])), np.max(np.abs(ref[2])))
    for name, a, b in zip("puvtq", got, ref):
        e = rel(a, b, suv if name in "uv" else None)
        assert e <= tol, "field %s: rel err %.3g > %g" % (name, e, tol)


# ---- 2-D shallow water: bit-exact ---------------------------------------------------------------------
def test_sw2d_operators_golden(backend):
    g = load_golden("sw2d_20x24")
    u, v, p, dx = g["u_0"], g["v_0"], g["p_0"], float(g["dx"])
    exact(matsuno_c_grid.advection_of_velocity_u(u, v, dx), g["adv_u"])
    exact(matsuno_c_grid.advection_of_velocity_v(u, v, dx), g["adv_v"])
    exact(matsuno_c_grid.geopotential_gradient_u(p, dx), g["grad_u"])
    exact(matsuno_c_grid.geopotential_gradient_v(p, dx), g["grad_v"])
    exact(matsuno_c_grid.advection_of_geopotential(u, v, p, dx), g["adv_p"])
    assert matsuno_c_grid.courant_number(p, u, dx, float(g["dt"])) == float(g["courant"])


def test_sw2d_matsuno_golden(backend):
    g = load_golden("sw2d_20x24")
    s = (g["u_0"], g["v_0"], g["p_0"])
    for a, k in zip(matsuno_c_grid.matsumo_scheme(*s, float(g["dx"]), float(g["dt"])), "uvp"):
        exact(a, g[k + "_1"])
    for a, k in zip(matsuno_c_grid.matsumo_scheme(*s, float(g["dx"]), float(g["dt"]), nsteps=50), "uvp"):
        exact(a, g[k + "_50"])


def test_sw2d_reference_main_case(backend):
    """matsuno_c_grid.py:146-157 at the stable bump: 64 x 64, dx = 300 km, dt = 300 s, 100 steps."""
    g = load_golden("sw2d_64x64_main")
    u = np.zeros((64, 64)); v = np.zeros((64, 64)); h = np.full((64, 64), 8000.0)
    u[32, 32] = 1.0
    for a, k in zip(matsuno_c_grid.matsumo_scheme(u, v, h, 300e3, 300.0, nsteps=100), "uvp"):
        exact(a, g[k + "_100"])


@pytest.mark.parametrize("H,W,n", [(80, 96, 3), (33, 70, 2), (1, 7, 2), (5, 1, 2)])
def test_sw2d_tiled_and_ragged_vs_oracle(backend, H, W, n):
    """Grids too large for the shared-memory-resident kernel take the tiled fused kernel; ragged and
    degenerate extents exercise the periodic wrap of the halo tiles."""
    rng = np.random.default_rng(H * 100 + W)
    u, v = rng.standard_normal((2, H, W))
    h = 8000.0 + 10 * rng.standard_normal((H, W))
    ref = (u, v, h)
    for _ in range(n):
        ref = O.matsumo_scheme(*ref, 300e3, 300.0)
    for a, b in zip(matsuno_c_grid.matsumo_scheme(u, v, h, 300e3, 300.0, nsteps=n), ref):
        exact(a, b)


# ---- coordinates: bit-exact shifts (reference KATs test_matsumo.py:9-21) ---------------------------------
def test_shift_directions(backend):
    p = np.full((3, 3), 1.0); p[1, 1] = 0
    assert coordinates.ipj(p)[1, 0] == 0          # test_matsumo.py:9-14
    assert coordinates.ijp(p)[0, 1] == 0          # test_matsumo.py:16-21
    assert matsuno_c_grid.geopotential_gradient_v(p, 1.0)[1, 1] == O.G     # test_matsumo.py:24-29
    rng = np.random.default_rng(0)
    q = rng.standard_normal((4, 5, 6))
    for name in ("ipj", "imj", "ijp", "ijm", "imjp", "kp", "km", "kph", "kmh", "iph", "imh", "jph", "jmh"):
        exact(getattr(coordinates_3d, name)(q), getattr(O, name)(q))
    exact(coordinates_3d.gradi(q, 3.0), O.gradi(q, 3.0))
    exact(coordinates_3d.gradj(q, 7.0), O.gradj(q, 7.0))
    q2 = q[0]
    for name in ("ipj", "imj", "ijp", "ijm", "imjp", "iph", "imh", "jph", "jmh"):
        exact(getattr(coordinates, name)(q2), getattr(O, name)(q2))
    q1 = q[0, 0]
    exact(coordinates_1d.ip(q1), O.ip1(q1)); exact(coordinates_1d.im(q1), O.im1(q1))
    exact(coordinates_1d.div(q1, 2.0), (q1 - O.im1(q1)) / 2.0)


# ---- viscosity / flux limiter / temperature ------------------------------------------------------------
def test_viscosity_golden(backend):
    g = load_golden("viscosity")
    exact(viscosity.finite_laplacian_2d(g["a"], 1.0), g["lap"])
    exact(viscosity.incompressible_viscosity_2d(g["b"], float(g["mu"]), 300e3), g["vis"])
    exact(viscosity.finite_laplacian_2d(g["c"], 3.5), g["lap2"])


def test_flux_limiter_golden(backend):
    g = load_golden("flux_limiter")
    assert flux_limiter.van_leer(1.0) == 1 and flux_limiter.van_leer(0.0) == 0      # flux_limiter.py:46-48
    exact(flux_limiter.van_leer(g["r"]), g["van_leer"])
    for tag in ("pos", "neg", "mix"):
        exact(flux_limiter.calc_r(g["q_" + tag]), g["r_" + tag])
        exact(flux_limiter.donor_cell_flux(g["q_" + tag], g["u_" + tag]), g["flux_" + tag])
    for tag in ("pos", "neg"):
        exact(flux_limiter.donor_cell_advection(g["q_" + tag], g["u_" + tag], 100.0, 1.0, nsteps=100), g["adv100_" + tag])
    exact(flux_limiter.donor_cell_advection(g["q_mix"], g["u_mix"], 100.0, 1.0), g["adv_mix"])


def test_temperature_roundtrip(backend):
    """temperature.py:31-41."""
    g = load_golden("temperature")
    th = temperature.to_potential_temp(np.array([O.standard_temperature]), np.array([O.standard_pressure]))
    assert rel(th, np.array([float(g["theta"])])) < 1e-15
    tt = temperature.to_true_temp(th, np.array([O.standard_pressure]))
    assert abs(tt[0] - O.standard_temperature) < 1e-7


# ---- 2-D primitive equations ---------------------------------------------------------------------------
def test_pe2d_golden(backend):
    g = load_golden("pe2d_24x36")
    s = tuple(g[k + "_0"] for k in "puvtq")
    dx, dt = float(g["dx"]), float(g["dt"])
    dut, dvt = no_limits_2d.advec_m(s[0], s[1], s[2], dx)
    exact(dut, g["dut"]); exact(dvt, g["dvt"])                                      # + - * / only
    for a, k in zip(no_limits_2d.pgf(s[0], s[3], dx), ("pgu", "pgv")):
        assert rel(a, g[k]) < TOL_CALL
    for n in (1, 20):
        out = no_limits_2d.matsuno_timestep(*s, dt, dx, nsteps=n)
        for a, k in zip(out, "puvtq"):
            assert rel(a, g["%s_%d" % (k, n)]) < TOL_RUN, k
        exact(out[4], g["q_%d" % n])                                                # q passes through


def test_pe2d_half_step_vs_oracle(backend):
    rng = np.random.default_rng(5)
    H, W = 10, 14
    p = 101325.0 + 50 * rng.standard_normal((H, W))
    u, v = rng.standard_normal((2, H, W))
    t = 280.0 + rng.standard_normal((H, W))
    q = rng.random((H, W))
    sp, su, sv, st = p + rng.standard_normal((H, W)), u * 1.1, v * 0.9, t + 0.1
    ref = O.pe2d_half_timestep(p, u, v, t, q, sp, su, sv, st, q, 0.1, 100.0)
    got = no_limits_2d.half_timestep(p, u, v, t, q, sp, su, sv, st, q, 0.1, 100.0)
    for a, b, k in zip(got, ref, "puvtq"):
        assert rel(a, b) < TOL_CALL, k


# ---- matsumo_temp --------------------------------------------------------------------------
def test_matsumo_temp_golden(backend):
    g = load_golden("matsumo_temp_12x12")
    s = tuple(g[k + "_0"] for k in "uvpt")
    for n in (1, 10):
        out = matsumo_temp.matsumo_scheme(*s, float(g["dx"]), float(g["dt"]), nsteps=n)
        suv = max(np.max(np.abs(g["u_%d" % n])), np.max(np.abs(g["v_%d" % n])))
        for a, k in zip(out, "uvpt"):
            assert rel(a, g["%s_%d" % (k, n)], suv if k in "uv" else None) < TOL_RUN, k


# ---- phi_port -----------------------------------------------------------------------------------------------
def test_phi_port_golden(backend):
    g = load_golden("phi_port_24x36x9")
    geom = geometry.gen_geometry(24, 36, 9)
    phi = phi_port.PGF(np.transpose(g["t"]), np.transpose(g["p"]), geom)
    assert phi.shape == (36, 24, 9)
    assert rel(np.transpose(phi), g["phi"]) < TOL_CALL
    geom.heightmap = g["heightmap2"]
    phi = np.transpose(phi_port.PGF(np.transpose(g["t2"]), np.transpose(g["p2"]), geom))
    assert rel(phi, g["phi2"]) < TOL_CALL
    assert np.all(phi[:, :, 1:] == 0)                  # IMAX = 1 quirk: only column 0 is computed


# ---- 2.5-D operators ---------------------------------------------------------------------------------------
def test_ops25_golden(backend):
    g = load_golden("ops25_24x36x9")
    geom = geometry.gen_geometry(24, 36, 9, sig_func=geometry.manabe_sig)
    p, u, v, t, q = (g[k] for k in "puvtq")
    exact(dynamics.calc_pu(p, u), g["pu"]); exact(dynamics.calc_pv(p, v), g["pv"])
    exact(dynamics.un_pu(g["pu"], p), g["un_pu"]); exact(dynamics.un_pv(g["pv"], p), g["un_pv"])
    pit, sd = dynamics.aflux(g["pu"], g["pv"], geom)
    exact(pit, g["pit"]); exact(sd, g["sd"])                                        # sums in the reference's k order
    dut, dvt = dynamics.advec_m_pu(p, u, v, g["pu"], g["pv"], geom)
    exact(dut, g["dut"]); exact(dvt, g["dvt"])
    exact(dynamics.advec_t(g["pu"], g["pv"], t, geom), g["advec_t"])
    exact(dynamics.advec_sig(g["sd"], t, geom), g["advec_sig"])
    assert rel(dynamics.compute_geopotential(p, t, geom), g["phi"]) < TOL_CALL
    pgu, pgv, phiu, phiv = dynamics.pgf(p, t, geom)
    assert rel(pgu, g["pgu"]) < TOL_CALL and rel(pgv, g["pgv"]) < TOL_CALL
    phimax, pmax = np.max(np.abs(g["phi"])), np.max(p)
    assert rel(phiu, g["phiu"], pmax * phimax / np.min(geom.dx_j)) < TOL_CALL
    assert rel(phiv, g["phiv"], pmax * phimax / geom.dy) < TOL_CALL
    assert rel(temperature.to_true_temp(t, p * geom.sig + geom.ptop), g["true_temp"]) < TOL_CALL
    assert rel(low_pass.arakawa_1977(g["pu"], geom), g["filt3d"]) < TOL_CALL
    f2 = low_pass.arakawa_1977(p, geom)
    assert f2.shape == (1, 24, 36) and rel(f2, g["filt2d"]) < TOL_CALL              # 2-D in -> (1, H, W) out
    assert rel(low_pass.avrx(p, geom), g["avrx2d"]) < TOL_CALL
    hs = dynamics.half_timestep(p, u, v, t, q, p, u, v, t, q, float(g["dt"]), geom)
    check_state(hs, tuple(g["hs_" + k] for k in "puvtq"), 1e-12)


@pytest.mark.parametrize("W", [8, 16, 36, 72, 98, 288, 1440])
def test_polar_filter_sizes_vs_oracle(backend, W):
    """Mixed-radix rows: 2^k, 2^2 3^2, 2^3 3^2, 2 7^2 (generic radix), 2^5 3^2, 2^5 3^2 5."""
    H, L = 6, 3
    geom = geometry.gen_geometry(H, W, L)
    q = np.random.default_rng(W).standard_normal((L, H, W))
    ref = O.arakawa_1977(q, O.gen_geometry(H, W, L))
    assert rel(low_pass.arakawa_1977(q, geom), ref) < TOL_CALL


def test_polar_filter_single_column_is_identity(backend):
    geom = geometry.gen_geometry(4, 1, 2)
    q = np.arange(8.0).reshape(2, 4, 1)
    exact(low_pass.arakawa_1977(q, geom), q)            # low_pass.py:58-59


# ---- 2.5-D N-step runs ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,H,W,L", [("run25_8x8x3", 8, 8, 3), ("run25_24x36x9", 24, 36, 9),
                                        ("run25_24x36x9_refic", 24, 36, 9), ("run25_24x36x9_ptop", 24, 36, 9),
                                        ("run25_1x16x17_mountain", 1, 16, 17), ("run25_46x72x9", 46, 72, 9)])
def test_run25_golden(backend, name, H, W, L):
    g = load_golden(name)
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    geom.ptop = float(g["ptop"]); geom.heightmap = g["heightmap"]
    s = tuple(g[k + "_0"] for k in "puvtq")
    n, dt = int(g["nsteps"]), float(g["dt"])
    snaps = sorted(int(k[2:]) for k in g if k.startswith("p_") and k != "p_0")
    st = dynamics.Stepper(geom, *s)
    done = 0
    for i in snaps:
        st.step(dt, i - done)
        done = i
        check_state(st.download(), tuple(g["%s_%d" % (k, i)] for k in "puvtq"), TOL_RUN)
    assert done == n
    one = dynamics.matsuno_timestep(*s, dt, geom)              # the drop-in per-step call
    check_state(one, tuple(g["%s_1" % k] for k in "puvtq") if "p_1" in g else O.matsuno_timestep(*s, dt, _ogeom(g, H, W, L)), 1e-12)


@pytest.mark.parametrize("H,W,L,dt,n", [(10, 64, 9, 300.0, 3), (7, 32, 3, 300.0, 2), (12, 96, 9, 200.0, 2),
                                        (2, 32, 3, 100.0, 2), (3, 32, 9, 100.0, 1)])
@pytest.mark.parametrize("knob4", [2, 0])
def test_run25_tiled_update_vs_oracle(backend, H, W, L, dt, n, knob4):
    """Widths that are multiples of 32 take the shared-memory-tiled update kernel (asynchronous copies, halo ring with
    periodic wrap); H not a multiple of the tile height leaves a partial tile.  Single grids this narrow would take the
    one-thread-per-cell update by default (knob4 = 0); knob 4 = 2 keeps them on the tiled kernel, which is what the
    wide grids (1440) and the ensembles run."""
    from gcmiipy_b200 import _lib
    _lib.lib().gcm_tuning_knob(4, knob4)
    try:
        _run25_vs_oracle(H, W, L, dt, n)
    finally:
        _lib.lib().gcm_tuning_knob(4, 0)


def _run25_vs_oracle(H, W, L, dt, n):
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    hm = 50.0 * np.random.default_rng(H).random((H, W))
    geom.heightmap = hm; og.heightmap = hm
    s = O.synthetic_state(og, seed=H * W)
    ref = s
    for _ in range(n):
        ref = O.matsuno_timestep(*ref, dt, og)
    st = dynamics.Stepper(geom, *s)
    st.step(dt, n)
    check_state(st.download(), ref, TOL_RUN)


def _ogeom(g, H, W, L):
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    og.ptop = float(g["ptop"]); og.heightmap = g["heightmap"]
    return og


def test_run_model_golden(backend):
    """no_limits_2_5d.run_model(8, 8, 3, ...) through the STATS diagnostics (no_limits_2_5d.py:79-94, :220-236)."""
    g = load_golden("run_model_8x8x3")
    no_limits_2_5d.STATS.clear()
    seen = []
    p, u, v, t, q, ground, geom = no_limits_2_5d.run_model(8, 8, 3, 450.0, 5, lambda *s: seen.append(s[0].copy()))
    check_state((p, u, v, t, q), tuple(g[k] for k in "puvtq"), TOL_RUN)
    assert len(seen) == 5 and np.array_equal(seen[-1], p)
    assert rel(np.array(no_limits_2_5d.STATS["ke"]), g["energy"]) < 1e-12
    suv = max(np.max(np.abs(g["u"])), np.max(np.abs(g["v"])))      # u starts at 0: same scale as check_state
    assert rel(np.array(no_limits_2_5d.STATS["u_max"]), g["u_max"], suv) < TOL_RUN
    assert rel(np.array(no_limits_2_5d.STATS["v_min"]), g["v_min"], suv) < TOL_RUN
    ke = no_limits_2_5d.calc_energy(p, u, v, t, q, ground, geom)
    assert rel(np.array(ke), np.array(O.calc_energy(*(g[k] for k in "puvtq"), O.gen_geometry(8, 8, 3, sig_func=O.manabe_sig)))) < 1e-12


def test_full_timestep_and_boundary_callback(backend):
    geom = geometry.gen_geometry(8, 8, 3, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(8, 8, 3, sig_func=O.manabe_sig)
    s = O.synthetic_state(og, seed=3)
    bc = lambda p, u, v, t, q, dt, g: (p, u * 0.5, v, t, q)
    ref = O.matsuno_timestep(*s, 450.0, og, boundary_conditions=bc)
    got = dynamics.matsuno_timestep(*s, 450.0, geom, boundary_conditions=bc)
    check_state(got, ref, 1e-12)
    no_limits_2_5d.STATS.clear()
    out = no_limits_2_5d.full_timestep(*s, None, 450.0, 0.0, geom)
    check_state(out[:5], O.matsuno_timestep(*s, 450.0, og), 1e-12)
    assert len(no_limits_2_5d.STATS["ke"]) == 1


def test_ensemble_members_are_independent(backend):
    """Batched ensemble (BASELINE configs[3]): member m of a batched step == the same member stepped alone."""
    geom = geometry.gen_geometry(24, 36, 9, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(24, 36, 9, sig_func=O.manabe_sig)
    members = [O.synthetic_state(og, seed=1234 + m) for m in range(3)]
    batched = tuple(np.stack([m[f] for m in members]) for f in range(5))
    st = dynamics.Stepper(geom, *batched)
    st.step(450.0, 3)
    out = st.download()
    for m, s in enumerate(members):
        single = dynamics.Stepper(geom, *s)
        single.step(450.0, 3)
        for a, b in zip(out, single.download()):
            exact(a[m], b)
        ref = s
        for _ in range(3):
            ref = O.matsuno_timestep(*ref, 450.0, og)
        check_state(tuple(a[m] for a in out), ref, TOL_RUN)


@pytest.mark.parametrize("H,W,nm", [(24, 36, 5), (6, 72, 2), (5, 36, 4)])
def test_ensemble_on_36_wide_tiles_vs_oracle(backend, H, W, nm):
    """Ensembles of 36-wide members (BASELINE configs[3]: 1024 x 36 x 24 x 9; more than 128 / W members) take the
    shared-memory-tiled update with 36-wide tiles (`pe25f_update_tiled_kernel<L, 4, 3, 36>`: a tile spans the member's
    row, both periodic seams inside it; partial tiles in j); knob 4 = 6 keeps the direct-load column march.  Every
    member against the oracle, and both kernels against each other."""
    from gcmiipy_b200 import _lib
    geom = geometry.gen_geometry(H, W, 9, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, 9, sig_func=O.manabe_sig)
    hm = 60.0 * np.random.default_rng(W + H).random((H, W))
    geom.heightmap = hm; og.heightmap = hm
    members = []
    for m in range(nm):
        sm = list(O.synthetic_state(og, seed=900 + m))
        sm[2] = sm[2] + 0.2 * np.random.default_rng(m).standard_normal(sm[2].shape)    # v != 0 on the wall row too
        members.append(sm)
    batched = tuple(np.stack([m[f] for m in members]) for f in range(5))
    res = {}
    try:
        for mode in (0, 6):
            assert _lib.lib().gcm_tuning_knob(4, mode) == 0
            st = dynamics.Stepper(geom, *batched)
            st.step(200.0, 3)
            res[mode] = st.download()
    finally:
        _lib.lib().gcm_tuning_knob(4, 0)
    for a, b in zip(res[0], res[6]):
        assert rel(a, b) <= 1e-13
    for m, s in enumerate(members):
        ref = tuple(s)
        for _ in range(3):
            ref = O.matsuno_timestep(*ref, 200.0, og)
        check_state(tuple(a[m] for a in res[0]), ref, TOL_RUN)


def test_nonfinite_input_propagates_like_numpy(backend):
    """The reference has no error path: NaNs propagate and the caller polls (matsuno_c_grid.py:184-187)."""
    geom = geometry.gen_geometry(8, 8, 3, sig_func=geometry.manabe_sig)
    s = list(O.synthetic_state(O.gen_geometry(8, 8, 3, sig_func=O.manabe_sig)))
    s[1] = s[1].copy(); s[1][0, 2, 2] = np.nan
    out = dynamics.matsuno_timestep(*s, 450.0, geom)
    assert np.isnan(out[1]).any()


@pytest.mark.parametrize("H,W,L", [(64, 32, 3), (70, 36, 9), (24, 36, 9)])
def test_step_host_is_bit_identical_to_resident_step(backend, H, W, L):
    """Stepper.step_host -> gcm_pe25_matsuno_step_host: latitude blocks copied in, stepped (predictor rows recomputed
    across block edges, two-segment launches across the periodic edge) and copied out on three streams; on the
    emulator and for H < 64 the one-block form.  Must equal the device-resident step bit for bit."""
    import torch
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    s = O.synthetic_state(og, seed=H + W)
    ref = dynamics.Stepper(geom, *s)
    ref.step(100.0, 2)
    st = dynamics.Stepper(geom, *s)
    from gcmiipy_b200 import _lib
    _lib.lib().gcm_tuning_knob(6, 8 if H >= 64 else 0)       # small grids pipeline only on request
    hin = [torch.from_numpy(np.ascontiguousarray(a)).clone() for a in s]
    hout = [torch.empty_like(a) for a in hin]
    if torch.cuda.is_available() and backend == "gpu":
        hin = [a.pin_memory() for a in hin]
        hout = [a.pin_memory() for a in hout]
    for _ in range(2):
        st.step_host(hin, hout, 100.0, 1)
        if backend == "gpu":
            torch.cuda.synchronize()
        hin, hout = hout, hin
    for a, b in zip(hin, ref.download()):
        exact(a.numpy(), b)
    _lib.lib().gcm_tuning_knob(6, 0)
    for a, b in zip(st.download(), ref.download()):
        exact(a, b)


@pytest.mark.parametrize("H,W,L,nblk,nsteps", [(64, 32, 3, 8, 7), (70, 36, 9, 5, 6), (96, 64, 9, 3, 5), (24, 36, 9, 0, 3)])
def test_step_host_pipelined_across_steps(backend, H, W, L, nblk, nsteps):
    """Stepper.step_host(pipelined=True) -> gcm_pe25_matsuno_step_host_pipelined: consecutive steps overlap (rotated
    block order, per-block copy-out events, alternating device states, no join until host_join()); the host buffers
    ping-pong like a host-resident time loop.  Bit-identical to the device-resident run, every step.  On the emulator
    (no copy engines) the call reports GCM_EUNSUP and step_host takes the joining path: same result."""
    import torch
    from gcmiipy_b200 import _lib
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    s = O.synthetic_state(og, seed=H + 3 * W)
    ref = dynamics.Stepper(geom, *s)
    st = dynamics.Stepper(geom, *s)
    hin = [torch.from_numpy(np.ascontiguousarray(a)).clone() for a in s]
    hout = [torch.zeros_like(a) for a in hin]
    if torch.cuda.is_available() and backend == "gpu":
        hin = [a.pin_memory() for a in hin]
        hout = [a.pin_memory() for a in hout]
    try:
        _lib.lib().gcm_tuning_knob(6, nblk)
        for n in range(nsteps):
            st.step_host(hin, hout, 100.0, 1, pipelined=True)
            hin, hout = hout, hin
            if n in (1, nsteps - 1):                      # look at the host state in the middle and at the end
                st.host_join()
                if backend == "gpu":
                    torch.cuda.synchronize()
                ref.step(100.0, n + 1 - ref.nsteps_done)
                for a, b in zip(hin, ref.download()):
                    exact(a.numpy(), b)
        for a, b in zip(st.download(), ref.download()):
            exact(a, b)
    finally:
        _lib.lib().gcm_tuning_knob(6, 0)


# ---- full-size properties (BASELINE sizes, no CPU reference needed) ---------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("H,W,L,dt", [(180, 288, 9, 60.0), (720, 1440, 9, 10.0)])
def test_full_size_properties(H, W, L, dt):
    """At the benchmark grids (configs[2] and configs[4]) the oracle is too slow, so check what must hold at any size:
    (1) the flux form conserves the unweighted sum of p (sum_i of an i-difference and sum_j of a j-difference vanish
        on the periodic grid, dynamics.py:35-46, :194);
    (2) a zonal shift of the initial state by s columns shifts the result by s columns (periodic i, flat ground):
        exercises the tile halos, the filter rows and the wrap at every column;
    (3) the run stays finite and v stays zero on the wall row (dynamics.py:222)."""
    import torch
    from gcmiipy_b200 import _lib, synthetic
    assert torch.cuda.is_available()
    _lib._override_for_tests(None, None)
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    s = synthetic.synthetic_state(geom, seed=99)
    st = dynamics.Stepper(geom, *s)
    st.step(dt, 3)
    out = st.download()
    assert all(np.isfinite(a).all() for a in out)
    assert np.all(out[2][:, -1, :] == 0)
    assert abs(np.sum(out[0]) - np.sum(s[0])) <= 1e-12 * np.sum(np.abs(s[0]))
    shift = 37
    st2 = dynamics.Stepper(geom, *(np.roll(a, shift, axis=-1) for a in s))
    st2.step(dt, 3)
    check_state(tuple(np.roll(a, -shift, axis=-1) for a in st2.download()), out, 1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,L", [(46, 72, 9), (20, 64, 9), (24, 36, 9), (180, 288, 9)])
def test_programmatic_dependent_launch_is_bit_identical(H, W, L):
    """Tuning knob 9: the half-step kernels launched with the programmatic-stream-serialization attribute (0 / 1, the
    default) and with an early trigger of their dependents (2) must give the bits of the ordinary launches (3), single
    steps and the replayed step-pair graph alike (a kernel that read or overwrote a field before its predecessor had
    finished would show up here)."""
    import torch
    from gcmiipy_b200 import _lib
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    s = O.synthetic_state(og, seed=H + W)
    res = {}
    try:
        for mode in (3, 1, 2):
            assert _lib.lib().gcm_tuning_knob(9, mode) == 0
            st = dynamics.Stepper(geom, *s)
            for _ in range(3):
                st.step(60.0, 1)
            st.step(60.0, 12)
            torch.cuda.synchronize()
            res[mode] = st.download()
    finally:
        _lib.lib().gcm_tuning_knob(9, 0)
    for mode in (1, 2):
        for a, b in zip(res[3], res[mode]):
            exact(b, a)


@pytest.mark.parametrize("H,W,L", [(46, 72, 9), (24, 36, 9), (5, 18, 3), (1, 6, 9)])
def test_cell_update_kernel_matches_column_march(backend, H, W, L):
    """Narrow single grids take the one-thread-per-cell update (pe25f_update_cell_kernel); knob 4 = 2 forces the
    column march (pe25f_update_kernel).  Same expressions operand for operand: bit for bit on the emulator build, to
    the last bit or two on the GPU (FMA contraction is chosen per kernel by the compiler)."""
    from gcmiipy_b200 import _lib
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    hm = 30.0 * np.random.default_rng(W).random((H, W))
    geom.heightmap = hm; og.heightmap = hm
    s = O.synthetic_state(og, seed=3 * H + W)
    res = {}
    try:
        for mode in (0, 2):
            assert _lib.lib().gcm_tuning_knob(4, mode) == 0
            st = dynamics.Stepper(geom, *s)
            st.step(100.0, 3)
            res[mode] = st.download()
    finally:
        _lib.lib().gcm_tuning_knob(4, 0)
    for a, b in zip(res[0], res[2]):
        if backend == "emu":
            exact(a, b)            # no FMA contraction on the CPU build: bit for bit
        else:
            assert rel(a, b) <= 1e-14   # nvcc contracts a*b+c per kernel as it sees fit: last-bit differences
    ref = s
    for _ in range(3):
        ref = O.matsuno_timestep(*ref, 100.0, og)
    check_state(res[0], ref, TOL_RUN)


def test_single_column_like_standard_atmosphere_isa(backend):
    """standard_atmosphere_isa.py:15-40 feeds a 1 x 1 x 18 column to compute_geopotential: W = 1 skips the polar filter
    (low_pass.py:58-59), every i / j neighbour is the column itself."""
    geom = geometry.gen_geometry(1, 1, 18, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(1, 1, 18, sig_func=O.manabe_sig)
    p = np.full((1, 1), 101325.0)
    tt = np.linspace(288.0, 216.0, 18).reshape(18, 1, 1)
    t = O.to_potential_temp(tt, p * og.sig + og.ptop)
    assert rel(dynamics.compute_geopotential(p, t, geom), O.compute_geopotential(p, t, og)) <= TOL_CALL
    s = (p, np.zeros((18, 1, 1)), np.zeros((18, 1, 1)), t, np.full((18, 1, 1), 1e-3))
    got, ref = dynamics.matsuno_timestep(*s, 100.0, geom), O.matsuno_timestep(*s, 100.0, og)
    for a, b in zip(got, ref):
        assert np.max(np.abs(a - b)) <= 1e-11 * max(np.max(np.abs(b)), 1.0)


# ---- the benchmarked grids against the oracle (C3 / C5 had no reference-parity assertion on the GPU) ----
def _ws_field(geom, which, shape):
    """A work field the last half step left in the geometry's workspace (gcm_pe25_workspace_field)."""
    from gcmiipy_b200 import _lib
    from gcmiipy_b200.geometry import device_geom
    dg = device_geom(geom)
    off = _lib.lib().gcm_pe25_workspace_field(dg.handle, 1, which)
    assert off != 2 ** 64 - 1
    n = int(np.prod(shape))
    return dg._ws[off:off + n].reshape(shape).cpu().numpy()


@pytest.mark.parametrize("H,W", [(12, 288), (8, 1440), (6, 72), (6, 36), (6, 96)])
def test_fused_filter_plans_vs_oracle(backend, H, W):
    """The polar filter INSIDE the fused half step (pe25f_filter_kernel<9, MODE, PLAN>: compile-time plans 1440 =
    12.8.15, 288 = 12.8.3, 72 = 8.9, 36 = 12.3; 96 takes the runtime switch) against low_pass.arakawa_1977
    (low_pass.py:41-78) of the oracle: the filtered mass flux spu (dynamics.py:187-189) and the filtered pgfu + phiu
    (:202) are read back from the step's work fields, and the half step's result is compared too."""
    L = 9
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    s = O.synthetic_state(og, seed=W + H)
    rng = np.random.default_rng(W)
    s = (s[0], s[1] + rng.standard_normal(s[1].shape), s[2], s[3] + 0.3 * rng.standard_normal(s[3].shape), s[4])
    got = dynamics.half_timestep(*s, *s, 5.0, geom)
    spu = _ws_field(geom, 0, (L, H, W))
    pgfu = _ws_field(geom, 1, (L, H, W))
    spu_ref = O.arakawa_1977(O.calc_pu(s[0], s[1]), og)
    pgu, _pgv, phiu, _phiv = O.pgf(s[0], s[3], og)
    raw = pgu + phiu
    pgf_ref = O.arakawa_1977(raw, og)
    assert rel(spu, spu_ref) <= TOL_CALL
    # pgfu + phiu is a small difference of terms of size p * max|phi| / dx: tolerance relative to the unfiltered field
    assert rel(pgfu, pgf_ref, np.max(np.abs(raw))) <= 1e-12
    # the filter must have done something on these rows (a wrong multiplier order would pass a no-op check)
    assert rel(spu_ref, O.calc_pu(s[0], s[1])) > 1e-3
    check_state(got, O.half_timestep(*s, *s, 5.0, og), 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,L,dt,n", [(180, 288, 9, 60.0, 40), (720, 1440, 9, 10.0, 3)])
def test_benchmark_grids_vs_oracle(H, W, L, dt, n):
    """SURVEY 8(d) parity cases of the benchmarked grids: configs[2] (288 x 180 x 9, dt = 60 s, 40 steps) and
    configs[4] (1440 x 720 x 9, dt = 10 s, 3 steps) against np_oracle.matsuno_timestep (dynamics.py:230-237) at 1e-11."""
    import torch
    from gcmiipy_b200 import _lib
    assert torch.cuda.is_available()
    _lib._override_for_tests(None, None)
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    s = O.synthetic_state(og, seed=1234)
    st = dynamics.Stepper(geom, *s)
    st.step(dt, n)
    ref = s
    for _ in range(n):
        ref = O.matsuno_timestep(*ref, dt, og)
    check_state(st.download(), ref, TOL_RUN)


# ---- C-ABI error reporting: gcm_last_status and the asynchronous non-finite watch -----------------
def test_nonfinite_watch_and_last_status(backend):
    """matsuno_c_grid.py:184-187 polls np.isnan(u).any() on the host after every step; here the step kernels count
    non-finite writes on the device and the count is read by a stream-ordered 4-byte copy."""
    from gcmiipy_b200 import _lib, watch
    geom = geometry.gen_geometry(12, 32, 9, sig_func=geometry.manabe_sig)
    s = O.synthetic_state(O.gen_geometry(12, 32, 9, sig_func=O.manabe_sig), seed=5)
    w = watch.NonFiniteWatch()
    st = dynamics.Stepper(geom, *s)
    st.step(100.0, 2)
    w.poll()
    assert w.value(wait=True) == 0
    bad = [a.copy() for a in s]
    bad[3][4, 5, 6] = np.nan
    st = dynamics.Stepper(geom, *bad)
    st.step(100.0, 1)
    w.poll(reset=True)
    assert w.value(wait=True) > 0
    w.poll()
    assert w.value(wait=True) == 0          # the reset is ordered after the read
    # narrow grid (one thread per cell update) and the 2-D schemes feed the same counter
    g2 = geometry.gen_geometry(6, 12, 3)
    s2 = list(O.synthetic_state(O.gen_geometry(6, 12, 3), seed=6))
    s2[1] = s2[1].copy(); s2[1][1, 2, 3] = np.inf
    dynamics.matsuno_timestep(*s2, 10.0, g2)
    w.poll(reset=True)
    assert w.value(wait=True) > 0
    u = np.zeros((8, 8)); v = np.zeros((8, 8)); h = np.full((8, 8), 8000.0)
    matsuno_c_grid.matsumo_scheme(u, v, h, 300e3, 300.0)
    w.poll()
    assert w.value(wait=True) == 0
    u[3, 3] = np.nan
    matsuno_c_grid.matsumo_scheme(u, v, h, 300e3, 300.0)
    w.poll(reset=True)
    assert w.value(wait=True) > 0
    # gcm_last_status: the newest non-zero status of this thread, cleared on request
    watch.last_status(clear=True)
    assert watch.last_status() == 0
    assert _lib.lib().gcm_pe25_workspace_bytes(None, 1) == 0
    assert _lib.lib().gcm_geom_destroy(None) == 0
    assert watch.last_status() == 0
    assert _lib.lib().gcm_tuning_knob(99, 0) == -2
    assert watch.last_status() == -2
    assert _lib.lib().gcm_tuning_knob(0, 0) == 0
    assert watch.last_status(clear=True) == -2 and watch.last_status() == 0


@pytest.mark.parametrize("knobs", [{}, {4: 5}, {4: 5, 10: 3}, {4: 5, 13: 1}, {4: 5, 10: 3, 13: 1}, {4: 5, 11: 8}, {4: 5, 12: 1}])
@pytest.mark.parametrize("H,W", [(10, 64), (3, 32)])
def test_tma_update_variants_vs_oracle(backend, knobs, H, W):
    """pe25f_update_tma_kernel (warp-specialised: TMA box loads by a producer warp, stages handed over on mbarriers):
    knob 4 = 5 selects it (default: the LDGSTS tiled kernel); 4 layers in flight, 32 x 4 tiles, 3 CTAs per SM; knob 10 =
    3 layers in flight, knob 13 = 4 CTAs per SM, knob 11 = 8 tile rows, knob 12 = L2 promotion of the tensor maps.  Two tile
    columns / one (both seams in one CTA), partial tiles in j, the periodic wrap in i and j patched by the seam CTAs."""
    from gcmiipy_b200 import _lib
    geom = geometry.gen_geometry(H, W, 9, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, 9, sig_func=O.manabe_sig)
    hm = 50.0 * np.random.default_rng(H).random((H, W))
    geom.heightmap = hm; og.heightmap = hm
    s = list(O.synthetic_state(og, seed=H * W))
    s[2] = s[2] + 0.3 * np.random.default_rng(1).standard_normal(s[2].shape)   # v != 0 on the wall row: the j wrap matters
    try:
        for k, v in knobs.items():
            assert _lib.lib().gcm_tuning_knob(k, v) == 0
        st = dynamics.Stepper(geom, *s)
        st.step(20.0, 3)
        got = st.download()
    finally:
        for k in knobs:
            _lib.lib().gcm_tuning_knob(k, 0)
    ref = tuple(s)
    for _ in range(3):
        ref = O.matsuno_timestep(*ref, 20.0, og)
    check_state(got, ref, TOL_RUN)


@pytest.mark.parametrize("H,W,L", [(8, 32, 17), (6, 36, 18), (5, 64, 18), (6, 20, 17), (6, 24, 4)])
def test_fused_path_layer_counts_vs_oracle(backend, H, W, L):
    """The fused kernels are instantiated for 3, 9, 17 (test_geography.py:6-18 mountain strip) and 18 layers
    (standard_atmosphere_isa.py:15-40); any other count (here 4) takes the general four-kernel path.  Tiled (W % 32 == 0),
    direct-load and one-thread-per-cell update kernels, ptop != 0."""
    from gcmiipy_b200 import _lib
    from gcmiipy_b200.geometry import device_geom
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    geom.ptop = og.ptop = 1000.0 if W == 36 else 0.0
    hm = 80.0 * np.random.default_rng(L).random((H, W))
    geom.heightmap = hm; og.heightmap = hm
    s = O.synthetic_state(og, seed=L * W)
    st = dynamics.Stepper(geom, *s)
    st.step(30.0, 2)
    ref = s
    for _ in range(2):
        ref = O.matsuno_timestep(*ref, 30.0, og)
    check_state(st.download(), ref, TOL_RUN)


@pytest.mark.parametrize("H,W,nm", [(12, 288, 1), (8, 1440, 1), (24, 36, 5), (10, 72, 3)])
def test_aflux_fused_into_filter_vs_oracle(backend, H, W, nm):
    """Knob 16 = 1: aflux (dynamics.py:35-46) fused into the filter of the mass flux -- filter MODE 2 keeps the filtered
    packed row in shared memory and writes conv[k0] + conv[k0+1] per layer pair; the tiled update sums the pairs in a
    fixed order into pit and forms p_n = p - pit dt.  Against the oracle and against the separate aflux kernel."""
    from gcmiipy_b200 import _lib
    geom = geometry.gen_geometry(H, W, 9, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, 9, sig_func=O.manabe_sig)
    members = [O.synthetic_state(og, seed=40 + m) for m in range(nm)]
    batched = tuple(np.stack([m[f] for m in members]) for f in range(5)) if nm > 1 else members[0]
    res = {}
    try:
        for mode in (0, 1):
            assert _lib.lib().gcm_tuning_knob(16, mode) == 0
            st = dynamics.Stepper(geom, *batched)
            st.step(20.0, 3)
            res[mode] = st.download()
    finally:
        _lib.lib().gcm_tuning_knob(16, 0)
    for a, b in zip(res[0], res[1]):
        assert rel(a, b) <= 1e-13
    for m, s in enumerate(members):
        ref = tuple(s)
        for _ in range(3):
            ref = O.matsuno_timestep(*ref, 20.0, og)
        got = tuple(a[m] for a in res[1]) if nm > 1 else res[1]
        check_state(got, ref, TOL_RUN)


@pytest.mark.parametrize("H,W,L", [(20, 256, 9), (9, 320, 9), (3, 256, 3), (17, 288, 9), (5, 270, 18)])
def test_hydro_tile_kernel_matches_marching_kernel(backend, H, W, L):
    """pe25f_hydro_tile_kernel (the default from 256 columns up: one column per thread, RT + 1 = 9 warps per CTA, south neighbour
    through shared memory) against the marching warp kernel (knob 7 = 3): the same expressions operand for operand -- bit for bit on the emulator
    build, to the last bits on the GPU (FMA contraction is chosen per kernel) -- and against the oracle.  Partial tiles
    in j (H = 20, 9, 3, 17, 5), the periodic wrap of the south neighbour, ptop != 0, 18 layers (85 KB of shared memory)."""
    from gcmiipy_b200 import _lib
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    geom.ptop = og.ptop = 500.0 if W == 320 else 0.0
    hm = 40.0 * np.random.default_rng(H + W).random((H, W))
    geom.heightmap = hm; og.heightmap = hm
    s = O.synthetic_state(og, seed=H * 7 + W)
    res = {}
    try:
        for mode in (0, 3):
            assert _lib.lib().gcm_tuning_knob(7, mode) == 0
            st = dynamics.Stepper(geom, *s)
            st.step(15.0, 2)
            res[mode] = st.download()
    finally:
        _lib.lib().gcm_tuning_knob(7, 0)
    res[2] = res[0]
    for a, b in zip(res[3], res[2]):
        if backend == "emu":
            exact(a, b)
        else:
            assert rel(a, b) <= 1e-13
    ref = s
    for _ in range(2):
        ref = O.matsuno_timestep(*ref, 15.0, og)
    check_state(res[2], ref, TOL_RUN)
