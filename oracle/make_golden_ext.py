"""Generate tests/golden/run25_24x36x9_coriolis.npz: the reference's 2.5-D step with its dead Coriolis branch on.

TEST INFRASTRUCTURE ONLY.  Run in the build container:   python oracle/make_golden_ext.py

dynamics.py:82 reads `if False:` (the Coriolis block :83-95 is complete but switched off).  This script reads the
reference's dynamics.py, flips that one condition IN MEMORY, executes the text as a module next to the otherwise
unmodified reference (oracle/ref_loader.py stand-ins) and stores the outputs of `half_timestep` and of a few
`matsuno_timestep`s.  Nothing of the reference is copied into the repository; only the vectors are committed.
They pin oracle/np_oracle.py: coriolis_terms / half_timestep_ext(coriolis=True).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import np_oracle as O          # only for the seeded synthetic inputs
import ref_loader
from make_golden import m, save


def main():
    const, geometry, dynamics = ref_loader.load("constants", "geometry", "dynamics")
    U = const.units
    src = open(os.path.join(ref_loader.REFERENCE_DIR, "dynamics.py")).read()
    dead = "    if False:\n        pu_at_pv"
    assert src.count(dead) == 1, "dynamics.py:82 is not where it was"
    mod = types.ModuleType("dynamics_coriolis_on")
    mod.__file__ = "<dynamics.py with line 82 switched on>"
    exec(compile(src.replace(dead, "    if True:\n        pu_at_pv"), mod.__file__, "exec"), mod.__dict__)

    H, W, L, dt = 24, 36, 9, 450.0
    with ref_loader.quiet():
        g = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    go = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    s = O.synthetic_state(go, seed=1234)
    Q = lambda a, unit: np.array(a, dtype=float) * unit
    st = (Q(s[0], U.Pa), Q(s[1], U.m / U.s), Q(s[2], U.m / U.s), Q(s[3], U.K), Q(s[4], U.dimensionless))
    out = {}
    with ref_loader.quiet():
        pu, pv = mod.calc_pu(st[0], st[1]), mod.calc_pv(st[0], st[2])
        dut, dvt = mod.advec_m_pu(st[0], st[1], st[2], pu, pv, g)
        dut0, dvt0 = dynamics.advec_m_pu(st[0], st[1], st[2], pu, pv, g)
        hs = mod.half_timestep(*st, *st, dt * U.s, g)
        cur = st
        for i in range(1, 11):
            cur = mod.matsuno_timestep(*cur, dt * U.s, g)
            if i in (1, 10):
                for nm, a in zip("puvtq", cur):
                    out["%s_%d" % (nm, i)] = a
    assert all(np.isfinite(m(a)).all() for a in out.values())
    assert np.max(np.abs(m(dut) - m(dut0))) > 0
    save("run25_24x36x9_coriolis", p_0=s[0], u_0=s[1], v_0=s[2], t_0=s[3], q_0=s[4], dt=dt, nsteps=10,
         pu=pu, pv=pv, dut=dut, dvt=dvt, cor_u=m(dut) - m(dut0), cor_v=m(dvt) - m(dvt0),
         hs_p=hs[0], hs_u=hs[1], hs_v=hs[2], hs_t=hs[3], hs_q=hs[4], **out)


def barometric():
    """geometry.pressure_from_heightmap (geometry.py:185-231) on a few heights."""
    const, geometry = ref_loader.load("constants", "geometry")
    U = const.units
    h = np.array([[0.0, 100.0, 1000.0], [2500.0, 5000.0, 8848.0]])
    with ref_loader.quiet():
        p = geometry.pressure_from_heightmap(h * U.m, 101325.0 * U.Pa, 288.15 * U.K)
    save("barometric", height=h, p0=101325.0, t0=288.15, p=p)


if __name__ == "__main__":
    barometric()
    main()
