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


@pytest.mark.parametrize("seed", range(16))
def test_random_band_split_is_bit_identical(emu, seed):
    """Every band of a random split, halo rows taken from the whole-grid states: predictor and corrector of the band
    equal the whole-grid half steps BIT for bit on the rows the band owns (with and without opt-in terms)."""
    from gcmiipy_b200 import bands
    rng = np.random.default_rng(5000 + seed)
    world = int(rng.choice([2, 3, 4]))
    H = world * int(rng.integers(2, 6))
    W = int(rng.choice([8, 14, 20, 32, 36, 64]))
    L = int(rng.choice([3, 4, 9]))
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    og = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    geom.heightmap = 150.0 * rng.random((H, W))
    if rng.random() < 0.6:
        dynamics.configure(geom, coriolis=bool(rng.random() < 0.5), viscosity=float(rng.choice([0.0, 2.0e4])),
                           limit_q=True, limit_t=bool(rng.random() < 0.5))
    s = O.synthetic_state(og, seed=seed)
    dt = 60.0
    star = dynamics.half_timestep(*s, *s, dt, geom)
    new = dynamics.half_timestep(*s, *star, dt, geom)
    for rank in range(world):
        b = bands.BandStepper(geom, *s, rank=rank, world=world, native=False)
        hn, hs = b.halo_n, b.halo_s
        rows = np.arange(b.j0 - hn, b.j1 + hs) % H
        own = slice(hn, hn + b.j1 - b.j0)
        b._half(b.cur, b.cur, b.star, dt)
        for got, want in zip(b.star, star):
            assert np.array_equal(got.numpy()[..., own, :], want[..., b.j0:b.j1, :])
        for dst, src in zip(b.star, star):       # what the exchange of the star state would deliver
            dst.copy_(torch.from_numpy(np.ascontiguousarray(np.take(src, rows, axis=-2))))
        b._half(b.cur, b.star, b.nxt, dt)
        for got, want in zip(b.nxt, new):
            assert np.array_equal(got.numpy()[..., own, :], want[..., b.j0:b.j1, :])
