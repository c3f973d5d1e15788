"""Pin oracle/np_oracle.py: (a) the reference's own known-answer tests, (b) the golden vectors
produced by the UNMODIFIED reference (oracle/make_golden.py).  Bit-exact (np.array_equal)."""
import numpy as np
import pytest

import np_oracle as O
from conftest import load_golden


def eq(a, b):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b), "max |diff| = %g" % np.max(np.abs(a - b))


# ---- (a) reference KATs ---------------------------------------------------------------------------
def test_kat_ipj_direction():          # test_matsumo.py:9-14
    p = np.full((3, 3), 1.0); p[1, 1] = 0
    assert O.ipj(p)[1, 0] == 0


def test_kat_ijp_direction():          # test_matsumo.py:16-21
    p = np.full((3, 3), 1.0); p[1, 1] = 0
    assert O.ijp(p)[0, 1] == 0


def test_kat_pgf_v():                  # test_matsumo.py:24-29
    p = np.full((3, 3), 1.0); p[1, 1] = 0
    assert O.geopotential_gradient_v(p, 1.0)[1, 1] == O.G


def test_kat_van_leer():               # flux_limiter.py:46-48
    assert O.van_leer(1) == 1 and O.van_leer(0) == 0


def test_kat_temperature_roundtrip():  # temperature.py:31-41
    th = O.to_potential_temp(O.standard_temperature, O.standard_pressure)
    assert abs(O.to_true_temp(th, O.standard_pressure) - O.standard_temperature) < 1e-7
    g = load_golden("temperature")
    assert th == g["theta"] and O.to_true_temp(th, O.standard_pressure) == g["tt2"]


# ---- (b) golden vectors from the reference itself ----------------------------------------------
@pytest.mark.parametrize("H,W,L,sf", [(24, 36, 9, O.manabe_sig), (46, 72, 9, O.manabe_sig),
                                      (8, 8, 3, O.equal_sig), (1, 16, 17, O.manabe_sig)])
def test_geometry(H, W, L, sf):
    g = load_golden("geom_%dx%dx%d" % (H, W, L))
    geom = O.gen_geometry(H, W, L, sig_func=sf)
    for k in ("sige", "sigb", "sigt", "dsig", "sig", "dsigv", "dx_j", "dx_h", "dy", "area", "ptop", "lat", "long", "heightmap"):
        eq(getattr(geom, k), g[k])


def test_square_geometry():
    g = load_golden("geom_square_6x10x4")
    geom = O.gen_square_geometry(6, 10, 4, 300e3, 250e3, sig_func=O.manabe_sig)
    for k in ("sige", "dsig", "sig", "dx_j", "dx_h", "dy", "ptop", "heightmap"):
        eq(getattr(geom, k), g[k])


def test_initial_conditions():
    g = load_golden("ic_24x36x9")
    s = O.gen_initial_conditions(O.gen_geometry(24, 36, 9, sig_func=O.manabe_sig))
    for a, k in zip(s, "puvtq"):
        eq(a, g[k])


def test_ops25():
    g = load_golden("ops25_24x36x9")
    geom = O.gen_geometry(24, 36, 9, sig_func=O.manabe_sig)
    p, u, v, t, q = (g[k] for k in "puvtq")
    pu, pv = O.calc_pu(p, u), O.calc_pv(p, v)
    eq(pu, g["pu"]); eq(pv, g["pv"])
    pit, sd = O.aflux(pu, pv, geom)
    eq(pit, g["pit"]); eq(sd, g["sd"])
    dut, dvt = O.advec_m_pu(p, u, v, pu, pv, geom)
    eq(dut, g["dut"]); eq(dvt, g["dvt"])
    eq(O.compute_geopotential(p, t, geom), g["phi"])
    for a, k in zip(O.pgf(p, t, geom), ("pgu", "pgv", "phiu", "phiv")):
        eq(a, g[k])
    eq(O.advec_t(pu, pv, t, geom), g["advec_t"])
    eq(O.advec_sig(sd, t, geom), g["advec_sig"])
    eq(O.arakawa_1977(pu, geom), g["filt3d"])
    eq(O.arakawa_1977(p, geom), g["filt2d"])
    eq(O.avrx(p, geom), g["avrx2d"])
    eq(O.un_pu(pu, p), g["un_pu"]); eq(O.un_pv(pv, p), g["un_pv"])
    eq(O.to_true_temp(t, p * geom.sig + geom.ptop), g["true_temp"])
    hs = O.half_timestep(p, u, v, t, q, p, u, v, t, q, float(g["dt"]), geom)
    for a, k in zip(hs, "puvtq"):
        eq(a, g["hs_" + k])


@pytest.mark.parametrize("name,H,W,L", [("run25_8x8x3", 8, 8, 3), ("run25_24x36x9", 24, 36, 9),
                                        ("run25_24x36x9_refic", 24, 36, 9), ("run25_24x36x9_ptop", 24, 36, 9),
                                        ("run25_1x16x17_mountain", 1, 16, 17), ("run25_46x72x9", 46, 72, 9)])
def test_run25(name, H, W, L):
    g = load_golden(name)
    geom = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    geom.ptop = float(g["ptop"]); geom.heightmap = g["heightmap"]
    s = tuple(g[k + "_0"] for k in "puvtq")
    n = int(g["nsteps"])
    for i in range(1, n + 1):
        s = O.matsuno_timestep(*s, float(g["dt"]), geom)
        if ("p_%d" % i) in g:
            for a, k in zip(s, "puvtq"):
                eq(a, g["%s_%d" % (k, i)])


def test_run_model_8x8x3():
    g = load_golden("run_model_8x8x3")
    geom = O.gen_geometry(8, 8, 3, sig_func=O.manabe_sig)
    s = O.run_model_ic(geom)
    for i in range(int(g["nsteps"])):
        s = O.matsuno_timestep(*s, float(g["dt"]), geom)
        eq(np.array(O.calc_energy(*s, geom)), g["energy"][i])
        assert np.max(s[1]) == g["u_max"][i] and np.min(s[2]) == g["v_min"][i]
    for a, k in zip(s, "puvtq"):
        eq(a, g[k])


def test_sw2d():
    g = load_golden("sw2d_20x24")
    u, v, p = g["u_0"], g["v_0"], g["p_0"]
    dx, dt = float(g["dx"]), float(g["dt"])
    eq(O.advection_of_velocity_u(u, v, dx), g["adv_u"]); eq(O.advection_of_velocity_v(u, v, dx), g["adv_v"])
    eq(O.geopotential_gradient_u(p, dx), g["grad_u"]); eq(O.geopotential_gradient_v(p, dx), g["grad_v"])
    eq(O.advection_of_geopotential(u, v, p, dx), g["adv_p"])
    assert O.courant_number(p, u, dx, dt) == g["courant"]
    for i in range(1, 51):
        u, v, p = O.matsumo_scheme(u, v, p, dx, dt)
        if i in (1, 50):
            eq(u, g["u_%d" % i]); eq(v, g["v_%d" % i]); eq(p, g["p_%d" % i])


def test_sw2d_main_ic():
    g = load_golden("sw2d_64x64_main")
    u = np.zeros((64, 64)); v = np.zeros((64, 64)); p = np.full((64, 64), 8000.0); u[32, 32] = 1.0
    for i in range(100):
        u, v, p = O.matsumo_scheme(u, v, p, float(g["dx"]), float(g["dt"]))
    eq(u, g["u_100"]); eq(v, g["v_100"]); eq(p, g["p_100"])


def test_pe2d():
    g = load_golden("pe2d_24x36")
    s = tuple(g[k + "_0"] for k in "puvtq")
    dx, dt = float(g["dx"]), float(g["dt"])
    pgu, pgv = O.pe2d_pgf(s[0], s[3], dx)
    dut, dvt = O.pe2d_advec_m(s[0], s[1], s[2], dx)
    eq(pgu, g["pgu"]); eq(pgv, g["pgv"]); eq(dut, g["dut"]); eq(dvt, g["dvt"])
    for i in range(1, 21):
        s = O.pe2d_matsuno_timestep(*s, dt, dx)
        if i in (1, 20):
            for a, k in zip(s, "puvtq"):
                eq(a, g["%s_%d" % (k, i)])


def test_phi_port():
    g = load_golden("phi_port_24x36x9")
    geom = O.gen_geometry(24, 36, 9)
    eq(np.transpose(O.phi_port_PGF(np.transpose(g["t"]), np.transpose(g["p"]), geom)), g["phi"])
    geom.heightmap = g["heightmap2"]
    eq(np.transpose(O.phi_port_PGF(np.transpose(g["t2"]), np.transpose(g["p2"]), geom)), g["phi2"])


def test_viscosity():
    g = load_golden("viscosity")
    eq(O.finite_laplacian_2d(g["a"], 1.0), g["lap"])
    eq(O.incompressible_viscosity_2d(g["b"], O.mu_air, 300e3), g["vis"])
    assert O.mu_air == g["mu"]
    eq(O.finite_laplacian_2d(g["c"], 3.5), g["lap2"])


def test_flux_limiter():
    g = load_golden("flux_limiter")
    eq(O.van_leer(g["r"]), g["van_leer"])
    for tag in ("pos", "neg"):
        q, u = g["q_" + tag], g["u_" + tag]
        eq(O.calc_r(q), g["r_" + tag]); eq(O.donor_cell_flux(q, u), g["flux_" + tag])
        for i in range(100):
            q = O.donor_cell_advection(q, u, 100.0, 1.0)
        eq(q, g["adv100_" + tag])
    eq(O.calc_r(g["q_mix"]), g["r_mix"]); eq(O.donor_cell_flux(g["q_mix"], g["u_mix"]), g["flux_mix"])
    eq(O.donor_cell_advection(g["q_mix"], g["u_mix"], 100.0, 1.0), g["adv_mix"])


def test_matsumo_temp():
    g = load_golden("matsumo_temp_12x12")
    u, v, p, t = g["u_0"], g["v_0"], g["p_0"], g["t_0"]
    for i in range(1, 11):
        u, v, p, t = O.mt_matsumo_scheme(u, v, p, t, float(g["dx"]), float(g["dt"]))
        if i in (1, 10):
            eq(u, g["u_%d" % i]); eq(v, g["v_%d" % i]); eq(p, g["p_%d" % i]); eq(t, g["t_%d" % i])
