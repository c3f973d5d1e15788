"""Seeded sweep over random grid shapes on the emulator build of the kernel sources (CPU only): every combination of
row count, width (even, any prime factors), layer count (3 / 9 take the fused kernels, others the general path),
ptop, mountains and opt-in terms steps like the numpy oracle.  Catches index / wrap / dispatch mistakes that the
hand-picked shapes of test_parity.py would miss; the GPU parity tests proper stay in test_parity.py / test_extras.py."""
import numpy as np
import pytest
import torch

import np_oracle as O
from gcmiipy_b200 import _lib, dynamics, geometry
from test_parity import check_state

WIDTHS = [2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 30, 32, 36, 40, 64, 72, 96]


@pytest.fixture(scope="module")
def emu():
    from emu import emu_lib
    _lib._override_for_tests(emu_lib.load(), torch.device("cpu"))
    yield
    _lib._override_for_tests(None, None)


@pytest.mark.parametrize("seed", range(32))
def test_random_shape_steps_like_the_oracle(emu, seed):
    rng = np.random.default_rng(1000 + seed)
    H = int(rng.integers(1, 14))
    W = int(rng.choice(WIDTHS))
    L = int(rng.choice([1, 2, 3, 3, 4, 9, 9]))
    ptop = float(rng.choice([0.0, 0.0, 500.0]))
    nb = int(rng.choice([1, 1, 2]))
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    hm = 200.0 * rng.random((H, W)) * (rng.random() < 0.5)
    geom.ptop = og.ptop = ptop
    geom.heightmap = hm; og.heightmap = hm.copy()
    opt = None
    if rng.random() < 0.5:
        kw = dict(coriolis=bool(rng.random() < 0.5), viscosity=float(rng.choice([0.0, 3.0e4])),
                  limit_q=bool(rng.random() < 0.7), limit_t=bool(rng.random() < 0.5))
        if any(kw.values()):
            dynamics.configure(geom, **kw)
            opt = O.StepOptions(kw["coriolis"], kw["viscosity"], kw["limit_q"], kw["limit_t"])
    members = [O.synthetic_state(og, seed=7 * seed + m) for m in range(nb)]
    dt, n = 60.0, int(rng.integers(1, 4))
    refs = []
    for s in members:
        r = s
        for _ in range(n):
            r = O.matsuno_timestep_ext(*r, dt, og, opt) if opt else O.matsuno_timestep(*r, dt, og)
        refs.append(r)
    state = members[0] if nb == 1 else tuple(np.stack([m[f] for m in members]) for f in range(5))
    st = dynamics.Stepper(geom, *state)
    st.step(dt, n)
    got = st.download()
    for m, r in enumerate(refs):
        assert all(np.isfinite(a).all() for a in r), "unstable case: pick another seed"
        check_state(got if nb == 1 else tuple(a[m] for a in got), r, 1e-11)
