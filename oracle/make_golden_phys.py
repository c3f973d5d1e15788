"""Generate the fixtures of SURVEY 8(f3)/(f4): the Hansen topography and the grey-radiation column physics.

TEST INFRASTRUCTURE ONLY.  Run in the build container:   python oracle/make_golden_phys.py

  tests/golden/hansen_topography.npz      hansen_topography.calc_topography() (hansen_topography.py:80-96) of the
                                          UNMODIFIED reference: surface height [24, 36] in metres.  The same array is
                                          written to gcmiipy_b200/data/hansen_topography_8x10.npy -- DATA the product ships
                                          (the decoded map of Hansen et al. 1983, fig. on p. 611), not reference source.
  tests/golden/grey_radiation_24x36x9.npz grey_solar.zenith_angle / basic_grey_transmittances / basic_grey_radiation
                                          (grey_solar.py:49-68, :323-333, :358-563) and no_limits_2_5d.solar_timestep
                                          (:66-75) on seeded inputs at three model times; a run over the topography too.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import np_oracle as O          # only for the seeded synthetic inputs
import ref_loader
from make_golden import m, save


def main():
    const, geometry, dynamics, topo, gs, model = ref_loader.load("constants", "geometry", "dynamics", "hansen_topography",
                                                                 "grey_solar", "no_limits_2_5d")
    U = const.units
    with ref_loader.quiet():
        hmap = m(topo.calc_topography())
    assert hmap.shape == (24, 36)
    save("hansen_topography", heightmap=hmap)
    np.save(os.path.join(ROOT, "gcmiipy_b200", "data", "hansen_topography_8x10.npy"), hmap)

    H, W, L = 24, 36, 9
    with ref_loader.quiet():
        g = geometry.gen_geometry(H, W, L)                       # equal sigma layers (basic_grey_radiation needs them)
    go = O.gen_geometry(H, W, L)
    s = O.synthetic_state(go, seed=4321)
    rng = np.random.default_rng(11)
    p = s[0] * U.Pa
    t = s[3] * U.K
    gt = (285.0 + 10.0 * rng.random((H, W))) * U.K
    ground = model.GroundVars(gt, np.zeros((H, W)) * U.m, np.zeros((H, W)) * U.m, np.zeros((H, W)) * U.m)
    out = dict(p=p, t=t, gt=gt)
    for n, hours in enumerate((0.0, 7.5, 19.25)):
        utc = hours * U.hours
        with ref_loader.quiet():
            tp = p * g.sig + g.ptop
            tt = model.temperature.to_true_temp(t, tp)
            sza = gs.zenith_angle(g.long, g.lat, utc, g)
            lw, sw = gs.basic_grey_transmittances(0.1, 0.9, g)
            dTdt, dtg = gs.basic_grey_radiation(p, tp, tt, ground, 0.1, 0.9, 0.3, utc, g)
            t_n, g_n = model.solar_timestep(t, p, ground, 900.0 * U.s, utc, g)
        out.update({"utc_h_%d" % n: hours, "sza_%d" % n: sza, "dTdt_%d" % n: dTdt, "dtg_%d" % n: dtg,
                    "t_n_%d" % n: t_n, "gt_n_%d" % n: g_n.gt})
        if n == 0:
            out.update(tp=tp, tt=tt, lw_tr=lw, sw_tr=sw)
    # dynamics over the Hansen topography: 10 Matsuno steps on the 8 x 10 degree grid, 9 Manabe layers
    with ref_loader.quiet():
        g2 = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
        g2.heightmap = hmap * U.m
    go2 = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    s2 = O.synthetic_state(go2, seed=77)
    cur = (s2[0] * U.Pa, s2[1] * U.m / U.s, s2[2] * U.m / U.s, s2[3] * U.K, s2[4] * U.dimensionless)
    with ref_loader.quiet():
        for _ in range(10):
            cur = dynamics.matsuno_timestep(*cur, 300.0 * U.s, g2)
    out.update({"topo_%s_0" % k: a for k, a in zip("puvtq", s2)})
    out.update({"topo_%s_10" % k: a for k, a in zip("puvtq", cur)})
    assert all(np.isfinite(m(a)).all() for a in out.values())
    save("grey_radiation_24x36x9", **out)


if __name__ == "__main__":
    main()
