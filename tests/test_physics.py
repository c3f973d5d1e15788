"""SURVEY.md section 8 (f3) topography and (f4) column physics: hansen_topography.calc_topography wired into the step,
and the grey-radiation columns of no_limits_2_5d.solar_timestep (grey_solar.py:49-68, :323-333, :358-563).

Goldens: tests/golden/hansen_topography.npz and grey_radiation_24x36x9.npz, outputs of the UNMODIFIED reference
(oracle/make_golden_phys.py).  The oracle restatement is pinned to them bit for bit; the kernels (csrc/physics.cu,
-fmad=false, reference operation order) are bit-exact except for cos() of the hour angle and pow() of the Exner factor:
tolerance 1e-13 of max|ref| (written below)."""
import numpy as np
import pytest

import np_oracle as O
from conftest import load_golden
from gcmiipy_b200 import dynamics, geometry, grey_solar, hansen_topography, no_limits_2_5d

TOL = 1e-13
HOURS = 3600.0


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


# ---- oracle pinned to the reference (no GPU) ---------------------------------------------------------------------
def test_oracle_grey_radiation_is_bit_identical_to_the_reference():
    z = load_golden("grey_radiation_24x36x9")
    g = O.gen_geometry(24, 36, 9)
    lw, sw = O.basic_grey_transmittances(0.1, 0.9, g)
    assert np.array_equal(lw, z["lw_tr"]) and np.array_equal(sw, z["sw_tr"])
    tp = z["p"] * g.sig + g.ptop
    tt = O.to_true_temp(z["t"], tp)
    assert np.array_equal(tt, z["tt"])
    for n in range(3):
        utc = float(z["utc_h_%d" % n]) * HOURS
        assert np.array_equal(O.zenith_angle(g.long, g.lat, utc, g), z["sza_%d" % n])
        d, dg = O.basic_grey_radiation(z["p"], tp, tt, z["gt"], 0.1, 0.9, 0.3, utc, g)
        assert np.array_equal(d, z["dTdt_%d" % n]) and np.array_equal(dg, z["dtg_%d" % n])
        tn, gn = O.solar_timestep(z["t"], z["p"], z["gt"], 900.0, utc, g)
        assert np.array_equal(tn, z["t_n_%d" % n]) and np.array_equal(gn, z["gt_n_%d" % n])


def test_topography_data_is_the_reference_map():
    ref = load_golden("hansen_topography")["heightmap"]
    got = hansen_topography.calc_topography()
    assert got.shape == (24, 36) and got.dtype == np.float64
    assert np.array_equal(got, ref)
    # hansen_topography.py:84-93: ' ' -> 0, '0' -> 25, '1'..'9' -> 100..900, 'A'.. -> 1000.., '+' -> 4500
    vals = set(np.unique(got).tolist())
    assert vals <= ({0.0, 25.0, 4500.0} | {100.0 * k for k in range(1, 10)} | {1000.0 + 100.0 * k for k in range(26)})
    assert got.max() == 4500.0 and (got == 0).mean() > 0.4          # Tibet; oceans
    big = hansen_topography.regrid(48, 72)
    assert big.shape == (48, 72) and np.array_equal(big[::2, ::2], got)


# ---- kernels against the goldens ------------------------------------------------------------------------------
def test_zenith_and_transmittances_golden(backend):
    z = load_golden("grey_radiation_24x36x9")
    geom = geometry.gen_geometry(24, 36, 9)
    lw, sw = grey_solar.basic_grey_transmittances(0.1, 0.9, geom)
    assert np.array_equal(lw, z["lw_tr"]) and np.array_equal(sw, z["sw_tr"])
    for n in range(3):
        sza = grey_solar.zenith_angle(geom.long, geom.lat, float(z["utc_h_%d" % n]) * HOURS, geom)
        assert np.array_equal(sza, z["sza_%d" % n])


def test_basic_grey_radiation_golden(backend):
    z = load_golden("grey_radiation_24x36x9")
    geom = geometry.gen_geometry(24, 36, 9)
    g = no_limits_2_5d.GroundVars(z["gt"], None, None, None)
    for n in range(3):
        utc = float(z["utc_h_%d" % n]) * HOURS
        d, dg = grey_solar.basic_grey_radiation(z["p"], z["tp"], z["tt"], g, 0.1, 0.9, 0.3, utc, geom)
        assert rel(d, z["dTdt_%d" % n]) <= TOL and rel(dg, z["dtg_%d" % n]) <= TOL
    # night side: no short wave at all -> independent of the albedo, bit for bit
    utc = 12 * HOURS
    night = grey_solar.zenith_angle(geom.long, geom.lat, utc, geom) == 0
    a = grey_solar.basic_grey_radiation(z["p"], z["tp"], z["tt"], g, 0.1, 0.9, 0.3, utc, geom)
    b = grey_solar.basic_grey_radiation(z["p"], z["tp"], z["tt"], g, 0.1, 0.9, 0.9, utc, geom)
    assert night.any() and np.array_equal(a[1][night], b[1][night]) and np.array_equal(a[0][:, night], b[0][:, night])


def test_solar_timestep_golden(backend):
    z = load_golden("grey_radiation_24x36x9")
    geom = geometry.gen_geometry(24, 36, 9)
    zero = np.zeros((24, 36))
    g = no_limits_2_5d.GroundVars(z["gt"], zero, zero, zero)
    for n in range(3):
        utc = float(z["utc_h_%d" % n]) * HOURS
        t_n, g_n = no_limits_2_5d.solar_timestep(z["t"], z["p"], g, 900.0, utc, geom)
        assert rel(t_n, z["t_n_%d" % n]) <= TOL and rel(g_n.gt, z["gt_n_%d" % n]) <= TOL
        assert g_n.gw is zero and g_n.snow is zero


@pytest.mark.parametrize("H,W,L", [(5, 8, 3), (12, 20, 18)])
def test_solar_timestep_vs_oracle_other_shapes(backend, H, W, L):
    geom = geometry.gen_geometry(H, W, L)
    og = O.gen_geometry(H, W, L)
    s = O.synthetic_state(og, seed=H)
    gt = 280.0 + 15.0 * np.random.default_rng(W).random((H, W))
    utc = 5.3 * HOURS
    t_n, gt_n = grey_solar.solar_timestep(s[3], s[0], gt, 600.0, utc, geom)
    rt, rg = O.solar_timestep(s[3], s[0], gt, 600.0, utc, og)
    assert rel(t_n, rt) <= TOL and rel(gt_n, rg) <= TOL


def test_full_timestep_with_physics(backend):
    """full_timestep(physics=True): the dynamics step followed by the column physics the reference keeps below its
    early return (no_limits_2_5d.py:96-103)."""
    geom = geometry.gen_geometry(12, 12, 9)          # H == W: calc_energy's area broadcast (no_limits_2_5d.py:49)
    og = O.gen_geometry(12, 12, 9)
    s = O.synthetic_state(og, seed=3)
    gt = np.full((12, 12), 288.0)
    zero = np.zeros((12, 12))
    g = no_limits_2_5d.GroundVars(gt, zero, zero, zero)
    out = no_limits_2_5d.full_timestep(*s, g, 120.0, 2.0 * HOURS, geom, physics=True)
    ref = O.matsuno_timestep(*s, 120.0, og)
    rt, rg = O.solar_timestep(ref[3], ref[0], gt, 120.0, 2.0 * HOURS, og)
    assert rel(out[3], rt) <= 1e-11 and rel(out[5].gt, rg) <= 1e-11 and rel(out[0], ref[0]) <= 1e-11


# ---- dynamics over the Hansen topography -----------------------------------------------------------------------
def test_run_over_hansen_topography_golden(backend):
    """geom.heightmap = calc_topography(): 10 Matsuno steps of the reference over real mountains (4.5 km of Tibet)."""
    z = load_golden("grey_radiation_24x36x9")
    geom = geometry.gen_geometry(24, 36, 9, sig_func=geometry.manabe_sig)
    geom.heightmap = hansen_topography.calc_topography()
    s = tuple(z["topo_%s_0" % k] for k in "puvtq")
    st = dynamics.Stepper(geom, *s)
    st.step(300.0, 10)
    ref = tuple(z["topo_%s_10" % k] for k in "puvtq")
    suv = max(np.max(np.abs(ref[1])), np.max(np.abs(ref[2])))
    for name, a, b in zip("puvtq", st.download(), ref):
        scale = suv if name in "uv" else np.max(np.abs(b))
        assert np.max(np.abs(a - b)) / scale <= 1e-11, name


@pytest.mark.parametrize("world", [1, 3])
def test_band_solar_timestep_matches_whole_grid(backend, world):
    """f4 under the latitude-band layout ("trivially shardable"): every band runs the column physics on its
    stored rows with its own latitudes; owned rows equal the whole-grid result bit for bit."""
    import torch
    from gcmiipy_b200 import bands
    H, W, L = 24, 36, 9
    geom = geometry.gen_geometry(H, W, L)
    s = O.synthetic_state(O.gen_geometry(H, W, L), seed=21)
    gt = 280.0 + 12.0 * np.random.default_rng(2).random((H, W))
    utc = 9.25 * HOURS
    t_ref, gt_ref = grey_solar.solar_timestep(s[3], s[0], gt, 600.0, utc, geom)
    for rank in range(world):
        b = bands.BandStepper(geom, *s, rank=rank, world=world, native=False)
        rows = np.asarray(b.dg.rows)
        gt_n = b.solar_timestep(torch.from_numpy(np.ascontiguousarray(gt[rows])), 600.0, utc)
        lo, hi = b.dg.row_lo, b.dg.row_hi
        assert np.array_equal(b.cur[3].cpu().numpy()[:, lo:hi, :], t_ref[:, b.j0:b.j1, :])
        gt_n = gt_n.cpu().numpy() if hasattr(gt_n, "cpu") else gt_n
        assert np.array_equal(gt_n[lo:hi], gt_ref[b.j0:b.j1])
