"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not travel to the
GPU box):   python oracle/make_golden.py
The reference is imported through oracle/ref_loader.py (SI pint stand-in, no-op matplotlib,
np.float alias); inputs and outputs are stored as plain float64 SI magnitudes.  The fixtures pin
oracle/np_oracle.py (tests/test_oracle_pinning.py) and, through it and directly, the CUDA path.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import np_oracle as O          # only for the seeded synthetic inputs
import ref_loader

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def m(x):
    """SI magnitude as a plain ndarray copy."""
    return np.array(getattr(x, "m", x), dtype=float)


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: m(v) for k, v in arrs.items()})
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024.0))


def main():
    (const, geometry, dynamics, nl25, mcg, nl2d, low_pass, phi_port, viscosity, flux_limiter,
     temperature, mtemp, coords, coords3) = ref_loader.load(
        "constants", "geometry", "dynamics", "no_limits_2_5d", "matsuno_c_grid", "no_limits_2d",
        "low_pass", "phi_port", "viscosity", "flux_limiter", "temperature", "matsumo_temp",
        "coordinates", "coordinates_3d")
    U = const.units
    Q = lambda a, unit=None: (np.array(a, dtype=float) * (unit if unit is not None else U.dimensionless))
    quiet = ref_loader.quiet

    # ---- geometry tables -------------------------------------------------------------------
    for (H, W, L, sf) in [(24, 36, 9, "manabe_sig"), (46, 72, 9, "manabe_sig"), (8, 8, 3, "equal_sig"),
                          (1, 16, 17, "manabe_sig")]:
        with quiet():
            g = geometry.gen_geometry(H, W, L, sig_func=getattr(geometry, sf))
        save("geom_%dx%dx%d" % (H, W, L), sige=g.sige, sigb=g.sigb, sigt=g.sigt, dsig=g.dsig, sig=g.sig,
             dsigv=g.dsigv, dx_j=g.dx_j, dx_h=g.dx_h, dy=g.dy, area=g.area, ptop=g.ptop,
             lat=g.lat, long=g.long, heightmap=g.heightmap)
    with quiet():
        gs = geometry.gen_square_geometry(6, 10, 4, 300 * U.km, 250 * U.km, sig_func=geometry.manabe_sig)
    save("geom_square_6x10x4", sige=gs.sige, dsig=gs.dsig, sig=gs.sig, dx_j=gs.dx_j, dx_h=gs.dx_h, dy=gs.dy,
         ptop=gs.ptop, heightmap=gs.heightmap)

    # ---- 2.5-D: initial conditions, operators, half step, N-step runs ----------------------------
    def ref_geom(H, W, L, ptop=0.0, mountain=None):
        with quiet():
            g = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
        g.ptop = ptop * U.Pa
        if mountain is not None:
            g.heightmap.m[mountain[0], mountain[1]] = mountain[2]
        return g

    def ora_geom(H, W, L, ptop=0.0, mountain=None):
        g = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
        g.ptop = ptop
        if mountain is not None:
            g.heightmap[mountain[0], mountain[1]] = mountain[2]
        return g

    def wrap_state(s):
        p, u, v, t, q = s
        return (Q(p, U.Pa), Q(u, U.m / U.s), Q(v, U.m / U.s), Q(t, U.K), Q(q))

    g = ref_geom(24, 36, 9)
    with quiet():
        p, u, v, t, q, ground = nl25.gen_initial_conditions(g)
    save("ic_24x36x9", p=p, u=u, v=v, t=t, q=q)

    # operators on a perturbed state (all terms non-zero)
    s0 = O.synthetic_state(ora_geom(24, 36, 9), seed=1234)
    p, u, v, t, q = wrap_state(s0)
    with quiet():
        pu = dynamics.calc_pu(p, u)
        pv = dynamics.calc_pv(p, v)
        pit, sd = dynamics.aflux(pu, pv, g)
        dut, dvt = dynamics.advec_m_pu(p, u, v, pu, pv, g)
        phi = dynamics.compute_geopotential(p, t, g)
        pgu, pgv, phiu, phiv = dynamics.pgf(p, t, g)
        adt = dynamics.advec_t(pu, pv, t, g)
        ads = dynamics.advec_sig(sd, t, g)
        flt = low_pass.arakawa_1977(pu, g)
        flt2 = low_pass.arakawa_1977(p, g)
        avr = low_pass.avrx(p, g)
        unpu = dynamics.un_pu(pu, p)
        unpv = dynamics.un_pv(pv, p)
        tt = temperature.to_true_temp(t, p * g.sig + g.ptop)
        hs = dynamics.half_timestep(p, u, v, t, q, p, u, v, t, q, 450 * U.s, g)
    save("ops25_24x36x9", p=p, u=u, v=v, t=t, q=q, pu=pu, pv=pv, pit=pit, sd=sd, dut=dut, dvt=dvt, phi=phi,
         pgu=pgu, pgv=pgv, phiu=phiu, phiv=phiv, advec_t=adt, advec_sig=ads, filt3d=flt, filt2d=flt2,
         avrx2d=avr, un_pu=unpu, un_pv=unpv, true_temp=tt,
         hs_p=hs[0], hs_u=hs[1], hs_v=hs[2], hs_t=hs[3], hs_q=hs[4], dt=450.0)

    def run(g, s, dt, n, snaps=()):
        p, u, v, t, q = wrap_state(s)
        out = {}
        with quiet():
            for i in range(1, n + 1):
                p, u, v, t, q = dynamics.matsuno_timestep(p, u, v, t, q, dt * U.s, g)
                if i in snaps or i == n:
                    for nm, a in zip("puvtq", (p, u, v, t, q)):
                        out["%s_%d" % (nm, i)] = a
        assert all(np.isfinite(m(a)).all() for a in out.values())
        return out

    def run_case(name, H, W, L, dt, n, snaps=(), ptop=0.0, mountain=None, ic="synthetic", seed=1234):
        go = ora_geom(H, W, L, ptop, mountain)
        s = O.synthetic_state(go, seed=seed) if ic == "synthetic" else O.run_model_ic(go)
        out = run(ref_geom(H, W, L, ptop, mountain), s, dt, n, snaps)
        save(name, p_0=s[0], u_0=s[1], v_0=s[2], t_0=s[3], q_0=s[4], dt=dt, nsteps=n, ptop=ptop,
             heightmap=go.heightmap, **out)

    run_case("run25_8x8x3", 8, 8, 3, 450.0, 10, snaps=(1,))
    run_case("run25_24x36x9", 24, 36, 9, 450.0, 20, snaps=(1, 5))
    run_case("run25_24x36x9_refic", 24, 36, 9, 450.0, 10, snaps=(1,), ic="run_model")
    run_case("run25_24x36x9_ptop", 24, 36, 9, 450.0, 5, snaps=(1,), ptop=1000.0)
    run_case("run25_1x16x17_mountain", 1, 16, 17, 1800.0, 3, snaps=(1,), mountain=(0, 8, 1000.0), ic="run_model")
    run_case("run25_46x72x9", 46, 72, 9, 225.0, 20)

    # run_model(8, 8, 3, ...) through full_timestep, incl. calc_energy (no_limits_2_5d.py:79-94, :220-236)
    with quiet():
        nl25.STATS.clear()
        p, u, v, t, q, ground, g8 = nl25.run_model(8, 8, 3, 450 * U.s, 5, None)
        ke = np.array([[float(x) for x in e] for e in nl25.STATS["ke"]])
    save("run_model_8x8x3", p=p, u=u, v=v, t=t, q=q, energy=ke, u_max=np.array([float(x) for x in nl25.STATS["u_max"]]),
         v_min=np.array([float(x) for x in nl25.STATS["v_min"]]), dt=450.0, nsteps=5)

    # ---- 2-D shallow water (matsuno_c_grid.py) -------------------------------------------------
    rng = np.random.default_rng(7)
    H, W = 20, 24
    u0 = rng.standard_normal((H, W))
    v0 = rng.standard_normal((H, W))
    h0 = 8000.0 + 10 * rng.standard_normal((H, W))
    u, v, h = Q(u0, U.m / U.s), Q(v0, U.m / U.s), Q(h0, U.m)
    dx, dt = 300 * U.km, 300 * U.s
    out = dict(u_0=u0, v_0=v0, p_0=h0, dx=300e3, dt=300.0,
               adv_u=mcg.advection_of_velocity_u(u, v, dx), adv_v=mcg.advection_of_velocity_v(u, v, dx),
               grad_u=mcg.geopotential_gradient_u(h, dx), grad_v=mcg.geopotential_gradient_v(h, dx),
               adv_p=mcg.advection_of_geopotential(u, v, h, dx), courant=mcg.courant_number(h, u, dx, dt))
    for i in range(1, 51):
        u, v, h = mcg.matsumo_scheme(u, v, h, dx, dt)
        if i in (1, 50):
            out.update({"u_%d" % i: u, "v_%d" % i: v, "p_%d" % i: h})
    save("sw2d_20x24", **out)
    # the file's own IC (matsuno_c_grid.py:146-157) with the stable bump of SURVEY section 4: 64x64, u[32,32]=1
    u0 = np.zeros((64, 64)); v0 = np.zeros((64, 64)); h0 = np.full((64, 64), 8000.0)
    u0[32, 32] = 1.0
    u, v, h = Q(u0, U.m / U.s), Q(v0, U.m / U.s), Q(h0, U.m)
    for i in range(100):
        u, v, h = mcg.matsumo_scheme(u, v, h, dx, dt)
    save("sw2d_64x64_main", u_100=u, v_100=v, p_100=h, dx=300e3, dt=300.0)

    # ---- 2-D primitive equations (no_limits_2d.py), the file's own test IC (:156-166) ------------------
    H, W = 24, 36
    p = np.full((H, W), 1) * const.standard_pressure
    u = np.full((H, W), 1) * 1.0 * U.m / U.s
    v = np.full((H, W), 1) * .0 * U.m / U.s
    q = np.full((H, W), 1) * 0.1 * U.dimensionless
    t = temperature.to_potential_temp(np.full((H, W), 1) * const.standard_temperature, p)  # 0-d stand-in Quantity has .shape
    p[10, 10] *= 1.01
    u[0, 3] *= 200
    t[3, 3] *= 1.1
    dx, dt = 100 * U.m, .1 * U.s
    out = dict(p_0=p, u_0=u, v_0=v, t_0=t, q_0=q, dx=100.0, dt=0.1)
    pgu, pgv = nl2d.pgf(p, t, dx)
    dut, dvt = nl2d.advec_m(p, u, v, dx)
    out.update(pgu=pgu, pgv=pgv, dut=dut, dvt=dvt)
    for i in range(1, 21):
        p, u, v, t, q = nl2d.matsuno_timestep(p, u, v, t, q, dt, dx)
        if i in (1, 20):
            out.update({"p_%d" % i: p, "u_%d" % i: u, "v_%d" % i: v, "t_%d" % i: t, "q_%d" % i: q})
    save("pe2d_24x36", **out)

    # ---- phi_port.PGF on the test_phi_port.py fixture (:15-25) + a perturbed variant ---------------
    with quiet():
        g = geometry.gen_geometry(24, 36, 9)
    p = np.full((24, 36), 1) * const.standard_pressure - g.ptop
    t = np.full((9, 24, 36), 1) * temperature.to_potential_temp(np.full((24, 36), 1) * const.standard_temperature, p)
    phi = np.transpose(phi_port.PGF(np.transpose(t), np.transpose(p), g))
    rng = np.random.default_rng(11)
    p2 = p + Q(100 * rng.standard_normal((24, 36)), U.Pa)
    t2 = t + Q(rng.standard_normal((9, 24, 36)), U.K)
    g.heightmap.m[:] = 50 * rng.standard_normal((24, 36))
    phi2 = np.transpose(phi_port.PGF(np.transpose(t2), np.transpose(p2), g))
    save("phi_port_24x36x9", p=p, t=t, phi=phi, p2=p2, t2=t2, heightmap2=g.heightmap, phi2=phi2)

    # ---- viscosity.py on the test_viscosity.py fixtures (:9-19) --------------------------------------
    a = np.full((5, 5), const.standard_temperature, dtype=float) * U.kelvin
    a[0] = 100 * U.kelvin + const.standard_temperature
    lap = viscosity.finite_laplacian_2d(a, 1 * U.meter)
    b = np.zeros((5, 5)) * U.m / U.s
    b[2, 2] = 1 * U.m / U.s
    vis = viscosity.incompressible_viscosity_2d(b, const.mu_air, 300 * U.km)
    c = Q(rng.standard_normal((7, 9)))
    lap2 = viscosity.finite_laplacian_2d(c, 3.5 * U.m)
    save("viscosity", a=a, lap=lap, b=b, vis=vis, mu=const.mu_air, c=c, lap2=lap2)

    # ---- flux_limiter.py on its own fixtures (:74-79, :92-97) -----------------------------------------
    out = {}
    r = Q(np.array([-3.0, -1.0, -0.5, 0.0, 0.25, 1.0, 2.0, 1e6]))
    out["r"] = r
    out["van_leer"] = flux_limiter.van_leer(r)
    for tag, n, uval, lo, hi in (("pos", 16, 10.0, 4, 8), ("neg", 160, -10.0, 40, 80)):
        uu = np.full((n,), uval) * U.m / U.s
        qq = np.full((n,), 0.0) * U.gram / U.kg
        qq[lo:hi] = 1 * qq.u
        out["q_" + tag] = qq
        out["u_" + tag] = uu
        out["r_" + tag] = flux_limiter.calc_r(qq)
        out["flux_" + tag] = flux_limiter.donor_cell_flux(qq, uu)
        for i in range(100):
            qq = flux_limiter.donor_cell_advection(qq, uu, 100 * U.m, 1 * U.s)
        out["adv100_" + tag] = qq
    umix = Q(rng.standard_normal(32), U.m / U.s)
    qmix = Q(rng.standard_normal(32))
    out.update(u_mix=umix, q_mix=qmix, r_mix=flux_limiter.calc_r(qmix), flux_mix=flux_limiter.donor_cell_flux(qmix, umix),
               adv_mix=flux_limiter.donor_cell_advection(qmix, umix, 100 * U.m, 1 * U.s))
    save("flux_limiter", **out)

    # ---- temperature round trip (temperature.py:31-41) -----------------------------------------------
    th = temperature.to_potential_temp(float(const.standard_temperature), float(const.standard_pressure))
    save("temperature", tt=const.standard_temperature, p=const.standard_pressure, theta=th,
         tt2=temperature.to_true_temp(float(th), float(const.standard_pressure)))

    # ---- matsumo_temp.matsumo_scheme ------------------------------------------------------
    n = 12
    u, v, p, t = mtemp.gen_initial_conditions(n)
    p[n // 2, n // 2] += 100 * U.Pa
    u[3, 4] += 1 * U.m / U.s
    out = dict(u_0=u, v_0=v, p_0=p, t_0=t, dx=300e3, dt=300.0)
    for i in range(1, 11):
        u, v, p, t = mtemp.matsumo_scheme(u, v, p, t, 300 * U.km, 300 * U.s)
        if i in (1, 10):
            out.update({"u_%d" % i: u, "v_%d" % i: v, "p_%d" % i: p, "t_%d" % i: t})
    save("matsumo_temp_12x12", **out)


if __name__ == "__main__":
    main()
