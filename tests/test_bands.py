"""Latitude-band decomposition: an R-rank run must be BIT-IDENTICAL to the 1-rank run
(same kernels, decomposition-invariant arithmetic order).  World sizes 2 and 4 run here on the CPU with the
`gloo` backend and the emulator build of the kernels; the `gpu` cases run the same band code on a B200."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import np_oracle as O
from gcmiipy_b200 import bands, dynamics, geometry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(H=24, W=36, L=9, seed=1234):
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    geom.heightmap = 100.0 * np.random.default_rng(seed).random((H, W))       # rows differ: catches row mix-ups
    s = O.synthetic_state(O.gen_geometry(H, W, L, sig_func=O.manabe_sig), seed=seed)
    return geom, s


def test_single_rank_band_is_bit_identical_to_whole_grid(backend):
    geom, s = _case()
    whole = dynamics.Stepper(geom, *s)
    whole.step(450.0, 4)
    band = bands.BandStepper(geom, *s, rank=0, world=1)
    band.step(450.0, 4)
    for a, b in zip(band.gather(), whole.download()):
        assert np.array_equal(a, b)


@pytest.fixture(params=[0, 2], ids=["auto", "no-cell-kernel"])
def knob4(request, backend):
    """Narrow single grids take the one-thread-per-cell update by default; knob 4 = 2 keeps them on the tiled /
    column-march update kernels (what wide grids and ensembles run), so both are covered on bands."""
    from gcmiipy_b200 import _lib
    _lib.lib().gcm_tuning_knob(4, request.param)
    yield request.param
    _lib.lib().gcm_tuning_knob(4, 0)


@pytest.mark.parametrize("H,W,worlds", [(24, 36, (2, 3, 4)), (18, 32, (2, 3))])
def test_band_of_every_rank_matches_whole_grid_after_one_half_exchange(backend, knob4, H, W, worlds):
    """Bands stepped one by one in this process with halo rows taken from the whole-grid state.  W = 32 with 9 and 6
    owned rows: partial tiles of the shared-memory-tiled update kernel at the band's southern edge."""
    geom, s = _case(H=H, W=W)
    whole = dynamics.Stepper(geom, *s)
    whole.step(450.0, 1)
    ref = whole.download()
    star = dynamics.half_timestep(*s, *s, 450.0, geom)
    for world in worlds:
        for rank in range(world):
            b = bands.BandStepper(geom, *s, rank=rank, world=world, native=False)
            rows = np.arange(b.j0 - 1, b.j1 + 2) % H
            b._half(b.cur, b.cur, b.star, 450.0)
            for got, want in zip(b.star, star):
                got = got.cpu().numpy()
                assert np.array_equal(got[..., 1:-2, :], np.take(want, rows, axis=-2)[..., 1:-2, :])
            for dst, src in zip(b.star, star):       # halo rows of the star state from the whole-grid predictor
                dst.copy_(torch.from_numpy(np.ascontiguousarray(np.take(src, rows, axis=-2))))
            b._half(b.cur, b.star, b.nxt, 450.0)
            for got, want in zip(b.nxt, ref):
                assert np.array_equal(got.cpu().numpy()[..., 1:-2, :], want[..., b.j0:b.j1, :])


@pytest.mark.parametrize("native", [False, True])
@pytest.mark.parametrize("H,W", [(24, 36), (16, 32), (12, 14)])
def test_single_rank_band_with_options_is_bit_identical(backend, H, W, native):
    """The ring closed on the band itself, opt-in terms on (2 + 2 halo rows, two rows per exchange): narrow fused
    kernels (36), the 32-wide grid, the general 4-kernel path (14 has a factor 7); the torch.distributed schedule and
    the native loop of csrc/comm.cu."""
    geom, s = _case(H=H, W=W, L=9 if W != 14 else 4)
    dynamics.configure(geom, **OPTS)
    whole = dynamics.Stepper(geom, *s)
    whole.step(300.0, 3)
    band = bands.BandStepper(geom, *s, rank=0, world=1, native=native)
    assert (band.comm is not None) == native and (band.halo_n, band.halo_s) == (2, 2)
    band.step(300.0, 3)
    for a, b in zip(band.gather(), whole.download()):
        assert np.array_equal(a, b)
    with pytest.raises(ValueError):                      # the reference's halo widths are not enough
        geometry.device_geom(geom, band=(0, H // 2, 1, 2))


def test_band_diagnostics_single_rank(backend):
    geom, s = _case()
    band = bands.BandStepper(geom, *s, rank=0, world=1)
    band.step(450.0, 2)
    full = band.gather()
    d = band.diagnostics()
    assert d == {"u_max": np.max(full[1]), "u_min": np.min(full[1]), "v_max": np.max(full[2]),
                 "v_min": np.min(full[2]), "nonfinite": 0}


@pytest.mark.parametrize("W", [32, 288])
def test_band_with_tiled_update_is_bit_identical(backend, knob4, W):
    """W = 32: the band's interior rows run the shared-memory-tiled update kernel, the whole grid too.  W = 288: a
    compile-time FFT plan, so aflux is fused into the filter (per-pair partial sums of conv, summed by the update)."""
    from gcmiipy_b200 import _lib
    geom, s = _case(H=16, W=W)
    try:
        for fused in ((0, 1) if W == 288 else (0,)):
            _lib.lib().gcm_tuning_knob(16, fused)
            whole = dynamics.Stepper(geom, *s)
            whole.step(450.0, 2)
            band = bands.BandStepper(geom, *s, rank=0, world=1, native=True)
            band.step(450.0, 2)
            for a, b in zip(band.gather(), whole.download()):
                assert np.array_equal(a, b)
    finally:
        _lib.lib().gcm_tuning_knob(16, 0)


@pytest.mark.parametrize("wide", [True, False])
def test_native_band_loop_single_rank(backend, knob4, wide):
    """gcm_band_matsuno_step (csrc/comm.cu): the C++ loop with the ring closed on the band itself.  wide: 2 + 4 halo
    rows, one exchange per step, predictor recomputed on the rows across the band edges; else 1 + 2 halo rows, two
    exchanges, and on the GPU the overlapped schedule (side stream, interior rows first, two-segment launches)."""
    geom, s = _case()
    whole = dynamics.Stepper(geom, *s)
    band = bands.BandStepper(geom, *s, rank=0, world=1, native=True, wide_halo=wide)
    assert band.comm is not None and (band.halo_n, band.halo_s) == ((2, 4) if wide else (1, 2))
    assert band.overlap == 0          # default: exchange, then the step (the overlapped schedules measured slower)
    for mode in (1, 2, 0):            # 1 / 2: exchange beside the corrector's / the predictor's interior rows
        band.overlap = mode
        for n in (3, 2, 1):                                 # odd, even, single: both buffer parities, first / last step
            whole.step(450.0, n)
            band.step(450.0, n)
            for a, b in zip(band.gather(), whole.download()):
                assert np.array_equal(a, b)


def test_half_step_on_row_segments_equals_whole_band(backend):
    """gcm_pe25_half_step_rows: interior rows, then the rows next to the halos as one two-segment launch."""
    import ctypes
    from gcmiipy_b200 import _host, _lib
    from gcmiipy_b200.dynamics import _struct, _workspace
    geom, s = _case(H=24)
    b = bands.BandStepper(geom, *s, rank=1, world=3, native=False)
    rows = np.arange(b.j0 - 1, b.j1 + 2) % 24
    full = [torch.from_numpy(np.ascontiguousarray(np.take(x, rows, axis=-2))).to(b.cur[0].device) for x in s]
    for dst, src in zip(b.cur, full):
        dst.copy_(src)
    b._half(b.cur, b.cur, b.star, 450.0)
    lo, hi = b.dg.row_lo, b.dg.row_hi
    n = hi - lo
    out = [torch.zeros_like(x) for x in b.cur]
    ws, need = _workspace(b.dg, 1)
    sb, so = _struct(b.cur), _struct(out)
    seg = lambda *v: (ctypes.c_int * 4)(*v)
    for sr, su in ((seg(lo + 1, n - 2, 0, 0), seg(lo + 1, n - 3, 0, 0)), (seg(lo, 1, hi - 1, 2), seg(lo, 1, hi - 2, 2))):
        _lib.check(_lib.lib().gcm_pe25_half_step_rows(b.dg.handle, ctypes.byref(sb), ctypes.byref(sb), ctypes.byref(so),
                                                      450.0, 1, _host.ptr(ws), need, sr, su, _lib.stream()), "rows")
    for got, want in zip(out, b.star):
        assert np.array_equal(got.cpu().numpy()[..., lo:hi, :], want.cpu().numpy()[..., lo:hi, :])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


OPTS = dict(coriolis=True, viscosity=1.0e5, limit_q=True, limit_t=True)


def _worker(rank, world, port, nsteps, out, options=None):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["GCM_EMU_THREADS"] = "2"
    from emu import emu_lib
    from gcmiipy_b200 import _lib
    _lib._override_for_tests(emu_lib.load(), torch.device("cpu"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    geom, s = _case()
    if options:
        dynamics.configure(geom, **options)
    b = bands.BandStepper(geom, *s)
    assert (b.rank, b.world) == (rank, world)
    assert (b.halo_n, b.halo_s) == ((2, 2) if options else (1, 2))
    b.step(450.0, nsteps)
    full = b.gather()
    diag = b.diagnostics()                       # all-reduce over the ring: every rank gets the whole-grid values
    assert diag["nonfinite"] == 0
    assert diag["u_max"] == np.max(full[1]) and diag["u_min"] == np.min(full[1])
    assert diag["v_max"] == np.max(full[2]) and diag["v_min"] == np.min(full[2])
    if rank == 0:
        np.savez(out, *full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,options", [(2, None), (4, None), (2, OPTS), (3, OPTS)])
def test_gloo_ranks_bit_identical_to_single_process(world, options, tmp_path):
    """options: the opt-in terms (Coriolis, viscosity, flux-limited tracers) reach j - 2 ... j + 2, so the bands carry
    two halo rows on either side and exchange two rows each way; still bit-identical to the whole-grid run."""
    torch.set_num_threads(1)
    out = str(tmp_path / "bands.npz")
    mp.spawn(_worker, args=(world, _free_port(), 3, out, options), nprocs=world, join=True)
    from emu import emu_lib
    from gcmiipy_b200 import _lib
    _lib._override_for_tests(emu_lib.load(), torch.device("cpu"))
    try:
        geom, s = _case()
        if options:
            dynamics.configure(geom, **options)
        whole = dynamics.Stepper(geom, *s)
        whole.step(450.0, 3)
        with np.load(out) as z:
            for k, b in zip(sorted(z.files, key=lambda n: int(n.split("_")[1])), whole.download()):
                assert np.array_equal(z[k], b), k
    finally:
        _lib._override_for_tests(None, None)


def _subgroup_worker(rank, world, port, out):
    """World of 3, the bands live on the sub-group {1, 2} (global rank 0 is not a member): the 128-byte id made by the
    group's first rank must reach every member (dist.broadcast takes a GLOBAL source rank), and the band
    run on the sub-group must equal the whole grid."""
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["GCM_EMU_THREADS"] = "2"
    from emu import emu_lib
    from gcmiipy_b200 import _lib
    _lib._override_for_tests(emu_lib.load(), torch.device("cpu"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grp = dist.new_group([1, 2])
    if rank in (1, 2):
        geom, s = _case()
        b = bands.BandStepper(geom, *s, group=grp)
        assert (b.rank, b.world) == (rank - 1, 2)

        def make(buf):
            for n in range(128):
                buf[n] = (7 * n + 3) % 251
        got = bytes(b._share_id(make))
        assert got == bytes((7 * n + 3) % 251 for n in range(128)), "id did not reach group rank %d" % b.rank
        b.step(450.0, 2)
        full = b.gather()
        if b.rank == 0:
            np.savez(out, *full)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_subgroup_bands(tmp_path):
    torch.set_num_threads(1)
    out = str(tmp_path / "sub.npz")
    mp.spawn(_subgroup_worker, args=(3, _free_port(), out), nprocs=3, join=True)
    from emu import emu_lib
    from gcmiipy_b200 import _lib
    _lib._override_for_tests(emu_lib.load(), torch.device("cpu"))
    try:
        geom, s = _case()
        whole = dynamics.Stepper(geom, *s)
        whole.step(450.0, 2)
        with np.load(out) as z:
            for k, b in zip(sorted(z.files, key=lambda n: int(n.split("_")[1])), whole.download()):
                assert np.array_equal(z[k], b), k
    finally:
        _lib._override_for_tests(None, None)


@pytest.mark.parametrize("world,H,W", [(3, 24, 36), (2, 16, 32), (4, 32, 64), (2, 24, 288)])
def test_peer_mailbox_ring_same_process(backend, world, H, W):
    """The peer-memory halo exchange of csrc/comm.cu (push kernel: my boundary rows -> the neighbours' mailboxes +
    message flag; pull kernel: wait for both flags, mailbox -> halo rows; two alternating slots) with every rank of the
    ring in THIS process, stepped phase by phase (all ranks push before any pulls, so no kernel waits for a later
    launch).  Two exchanges per Matsuno step: both slots and the sequence numbers are exercised.  Bit-identical to the
    whole grid; no pull may time out."""
    import ctypes
    from gcmiipy_b200 import _lib
    from gcmiipy_b200.dynamics import _struct
    geom, s = _case(H=H, W=W)
    whole = dynamics.Stepper(geom, *s)
    whole.step(450.0, 3)
    lib = _lib.lib()
    ranks = [bands.BandStepper(geom, *s, rank=r, world=world, native=False) for r in range(world)]
    for b in ranks:
        h = ctypes.c_void_p()
        _lib.check(lib.gcm_comm_create(world, b.rank, None, ctypes.byref(h)), "gcm_comm_create")
        b.comm = h
        _lib.check(lib.gcm_comm_peer_setup(b.comm, b.dg.handle, (ctypes.c_ubyte * 64)()), "peer_setup")
    for b in ranks:
        _lib.check(lib.gcm_comm_peer_connect(b.comm, ranks[b.north].comm, ranks[b.south].comm, 1), "peer_connect")
        assert lib.gcm_comm_peer_status(b.comm, None) == 2

    def exchange(which):
        for fn in (lib.gcm_halo_exchange_begin, lib.gcm_halo_exchange_end):     # = gcm_band_halo_peer phases 1, 2
            for b in ranks:
                st = _struct(getattr(b, which))
                _lib.check(fn(b.dg.handle, b.comm, ctypes.byref(st), b.xn, b.xs, _lib.stream()), "gcm_halo_exchange")

    for _ in range(3):
        exchange("cur")
        for b in ranks:
            b._half(b.cur, b.cur, b.star, 450.0)
        exchange("star")
        for b in ranks:
            b._half(b.cur, b.star, b.nxt, 450.0)
            b.cur, b.nxt = b.nxt, b.cur
    full = [torch.cat([b.owned()[f] for b in ranks], dim=-2).cpu().numpy() for f in range(5)]
    for a, ref in zip(full, whole.download()):
        assert np.array_equal(a, ref)
    for b in ranks:
        assert b.peer_timeouts() == 0
        comm, b.comm = b.comm, None          # this test owns the communicators
        lib.gcm_comm_destroy(comm)
