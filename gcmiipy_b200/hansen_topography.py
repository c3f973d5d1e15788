"""Surface topography of the 8 x 10 degree grid, mirror of the reference `hansen_topography` module
(hansen_topography.py:80-96 `calc_topography`; the map of Hansen et al. 1983, p. 611).

The reference decodes an ASCII-art table at import time; the product ships the DECODED array as data
(`data/hansen_topography_8x10.npy`, written by oracle/make_golden_phys.py from the unmodified reference and pinned to
it by tests/test_physics.py) -- metres, [24, 36], row 0 = north like every field of the model.  `geom.heightmap =
calc_topography()` wires it into every kernel of the step (the geopotential at the ground, dynamics.py:133).
"""
import os

import numpy as np

height = 24
width = 36
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "hansen_topography_8x10.npy")


def calc_topography():
    """hansen_topography.py:80-96 -> surface height in metres, float64 [24, 36]."""
    a = np.load(_DATA)
    assert a.shape == (height, width)
    return np.array(a, dtype=np.float64)


def regrid(nrows, ncols):
    """The same topography on an nrows x ncols grid of the same extent (nearest cell; no reference counterpart: the
    reference only ever runs the 24 x 36 grid with it)."""
    a = calc_topography()
    j = np.minimum((np.arange(nrows) + 0.5) * height / nrows, height - 1).astype(int)
    i = np.minimum((np.arange(ncols) + 0.5) * width / ncols, width - 1).astype(int)
    return a[np.ix_(j, i)]
