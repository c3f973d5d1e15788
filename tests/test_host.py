"""Host-side logic of the package (no device): geometry tables, filter table, initial conditions, unit
stripping -- against the golden vectors produced by the unmodified reference."""
import os
import sys

import numpy as np
import pytest

from conftest import load_golden
from gcmiipy_b200 import _host, geometry, humidity, no_limits_2_5d, synthetic
import np_oracle as O


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and np.array_equal(a, b)


@pytest.mark.parametrize("H,W,L,sf", [(24, 36, 9, "manabe_sig"), (46, 72, 9, "manabe_sig"), (8, 8, 3, "equal_sig"),
                                      (1, 16, 17, "manabe_sig")])
def test_gen_geometry_matches_reference(H, W, L, sf):
    g = load_golden("geom_%dx%dx%d" % (H, W, L))
    geom = geometry.gen_geometry(H, W, L, sig_func=getattr(geometry, sf))
    for k in ("sige", "sigb", "sigt", "dsig", "sig", "dsigv", "dx_j", "dx_h", "dy", "area", "ptop", "lat", "long",
              "heightmap"):
        eq(getattr(geom, k), g[k])


def test_gen_square_geometry_matches_reference():
    g = load_golden("geom_square_6x10x4")
    geom = geometry.gen_square_geometry(6, 10, 4, 300e3, 250e3, sig_func=geometry.manabe_sig)
    for k in ("sige", "dsig", "sig", "dx_j", "dx_h", "dy", "ptop", "heightmap"):
        eq(getattr(geom, k), g[k])


def test_filter_table_matches_oracle():
    for H, W in ((24, 36), (46, 72), (8, 8)):
        geom = geometry.gen_geometry(H, W, 3)
        ref = O.polar_filter_table(O.gen_geometry(H, W, 3), W)
        eq(geometry.polar_filter_table(geom, W), ref.reshape(H, W // 2 + 1))


def test_initial_conditions_match_reference():
    g = load_golden("ic_24x36x9")
    s = no_limits_2_5d.gen_initial_conditions(geometry.gen_geometry(24, 36, 9, sig_func=geometry.manabe_sig))
    for a, k in zip(s[:5], "puvtq"):
        eq(a, g[k])
    assert s[5]._fields == ("gt", "gw", "snow", "ice")


def test_synthetic_state_equals_oracle_generator():
    a = synthetic.synthetic_state(geometry.gen_geometry(24, 36, 9, sig_func=geometry.manabe_sig))
    b = O.synthetic_state(O.gen_geometry(24, 36, 9, sig_func=O.manabe_sig))
    for x, y in zip(a, b):
        eq(x, y)


def test_unit_stripping_with_a_pint_like_quantity():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "shims"))
    import pint
    U = pint.UnitRegistry()
    dx = 300 * U.km
    assert _host.scalar(dx) == 300e3
    a = np.arange(6.0).reshape(2, 3) * U.hPa
    eq(_host.magnitude(a), np.arange(6.0).reshape(2, 3) * 100.0)
    fam = _host.Family(a)
    assert fam.quantity is not None and not fam.torch
