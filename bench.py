#!/usr/bin/env python
"""Benchmark of the Matsuno C-grid hot path (BASELINE.json: cell-updates/s, fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--repeats R] [--workload c5|c3|c2|c4|c1|c1big|p2d]
                    [--impl native|reference] [--options coriolis,limit_q,limit_t,viscosity=NU]

One "step" = one full Matsuno step (predictor + corrector, dynamics.py:230-237) of the 2.5-D model over the
whole grid.  Default workload: the 0.25 deg grid 1440 x 720 x 9 (BASELINE.json configs[4], the grid the metric
"cell-updates/s at 1/2/4/8 B200" is quoted on; 307 MB of state > 126 MB L2, so every step streams from HBM).
N > 1 (torchrun, one rank per GPU): latitude-band decomposition, strong scaling.
The timed call (K steps from the same initial state) is repeated R >= 5 times (more until >= 200 ms are timed);
`value` / `ms_per_step` are the MEDIAN repeat (the sustained figure), `best_ms_per_step` the fastest.  `state_sha256`
is the SHA-256 of the gathered state after K steps: identical lines at N = 1, 2, 4, 8 prove that the latitude-band
decomposition is bitwise invariant.  c1 / c1big (2-D shallow water, matsuno_c_grid.py:125-142, 48 B per cell-update)
and p2d (2-D primitive equations, no_limits_2d.py:129-131, 64 B) are single-GPU workloads.
Prints ONE JSON line (rank 0).  `--impl reference` times the reference's CPU algorithm (oracle/np_oracle.py,
numpy, all host cores as independent processes) on a bounded sample of the same workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (H, W, L, dt, members, description)
    "c5": (720, 1440, 9, 10.0, 1, "0.25deg 1440x720x9 2.5-D Matsuno step (BASELINE configs[4])"),
    "c3": (180, 288, 9, 60.0, 1, "1x1.25deg 288x180x9 2.5-D Matsuno step (BASELINE configs[2])"),
    "c2": (46, 72, 9, 225.0, 1, "GISS 4x5deg 72x46x9 2.5-D Matsuno step (BASELINE configs[1])"),
    "c4": (24, 36, 9, 450.0, 1024, "ensemble of 1024 x 8x10deg 36x24x9 runs (BASELINE configs[3])"),
    # tuning aid: the per-rank share of configs[4] on 8 GPUs as a stand-alone periodic grid (not a BASELINE config)
    # tuning aid: a quarter of configs[4]; on 2 GPUs every rank holds a 90-row band, as on 8 GPUs of the full grid
    "c5q": (180, 1440, 9, 10.0, 1, "1440x180x9: a quarter of configs[4] (tuning aid: 2 ranks = the band size of 8 ranks on the full grid)"),
    "c5b8": (90, 1440, 9, 10.0, 1, "1440x90x9: one of 8 latitude bands of configs[4], stepped alone (tuning aid)"),
}
# 2-D schemes: name: (kind, H, W, dt, dx, description)
WORKLOADS_2D = {
    "c1": ("sw2d", 64, 64, 300.0, 300e3, "2-D shallow-water Matsuno C-grid 64x64, dx = 300 km, dt = 300 s (BASELINE configs[0]; the "
           "reference's dt = 700 s diverges after 8 steps, SURVEY 8d)"),
    "c1dt700": ("sw2d", 64, 64, 700.0, 300e3, "2-D shallow-water Matsuno C-grid 64x64 at the reference's own dt = 700 s "
                "(matsuno_c_grid.py:146-157): unstable, reported for 8 steps only (SURVEY 8d)"),
    "c1big": ("sw2d", 8192, 8192, 300.0, 300e3, "2-D shallow-water Matsuno C-grid 8192x8192 (HBM-resident: 1.6 GB of state)"),
    "p2d": ("pe2d", 4096, 4096, 100.0, 300e3, "2-D primitive equations (no_limits_2d) 4096x4096"),
}
METRIC = "cell_updates_per_sec"
UNIT = "cell-updates/s"


def b_alg(L):
    """Algorithmic bytes per 3-D cell-update: the prognostic state read once and
    written once per Matsuno step: 4 three-D fields in + 4 out + the 2-D p in + out."""
    return 64.0 + 16.0 / L


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the newest committed `ncu --set full`
    summary under profiles/ (tools/ncu_summary.py) that lists it; None if no capture names the kernel."""
    import csv
    import glob
    base = kernel.split("<")[0]
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_full_summary.csv")), reverse=True):
        with open(path) as f:
            rows = [r for r in csv.DictReader(f) if r.get("kernel", "").startswith(base)]
        if rows:
            mb = [float(r["dram_read_MB"]) + float(r["dram_write_MB"]) for r in rows]
            return {"bytes_per_launch": 1e6 * sum(mb) / len(mb), "source": os.path.relpath(path, ROOT),
                    "launches_averaged": len(mb)}
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(n)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def parse_options(text):
    """--options coriolis,limit_q,limit_t,viscosity=1e5 -> kwargs of dynamics.configure (opt-in terms)."""
    kw = {}
    for item in filter(None, (text or "").split(",")):
        k, _, v = item.partition("=")
        if k not in ("coriolis", "limit_q", "limit_t", "viscosity"):
            raise SystemExit("unknown option %r" % k)
        kw[k] = float(v) if k == "viscosity" else True
    return kw


def cpu_oracle_step_rate(H, W, L, dt, nsteps, seed=1234, options=None):
    """Time the CPU restatement of the reference (oracle/np_oracle.py) on an H x W x L grid; returns seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import np_oracle as O
    geom = O.gen_geometry(H, W, L, sig_func=O.manabe_sig)
    s = O.synthetic_state(geom, seed=seed)
    opt = None
    if options:
        opt = O.StepOptions(options.get("coriolis", False), options.get("viscosity", 0.0), options.get("limit_q", False),
                            options.get("limit_t", False))
    t0 = time.perf_counter()
    for _ in range(nsteps):
        s = O.matsuno_timestep_ext(*s, dt, geom, opt) if opt else O.matsuno_timestep(*s, dt, geom)
    return time.perf_counter() - t0


def _ref_worker(args):
    H, W, L, dt, warmup, steps, options = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    if warmup:
        cpu_oracle_step_rate(H, W, L, dt, warmup, options=options)
    return cpu_oracle_step_rate(H, W, L, dt, steps, options=options)


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (numpy; the reference is pure Python, so the arm is
    the pinned oracle port) on all host cores as independent single-threaded processes, each stepping a
    bounded sample (a 1440-wide grid of fewer rows: the per-cell arithmetic is identical)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    H, W, L, dt, members, desc = WORKLOADS[args.workload]
    options = parse_options(args.options)
    if options:
        desc += " + opt-in terms"
    cores = os.cpu_count() or 1
    full_step_s = 1.7e-6 * H * W * L * (members if members > 1 else 1)       # ~0.6 M cell-updates/s/core
    budget = 120.0
    frac = min(1.0, budget / max(1e-9, full_step_s * (args.steps + args.warmup)))
    if members > 1:
        rows, nmem = H, max(1, int(members * frac))
        sample = "%d of %d members per process" % (nmem, members)
        cells = H * W * L * nmem
        work = (H, W, L, dt, args.warmup * nmem, args.steps * nmem, options)
    else:
        rows = max(8, int(H * frac) // 2 * 2)
        sample = "%d x %d x %d rows-subset grid per process (of %d rows)" % (W, rows, L, H)
        cells = rows * W * L
        work = (rows, W, L, dt, args.warmup, args.steps, options)
    with mp.get_context("fork").Pool(cores) as pool:
        times = pool.map(_ref_worker, [work] * cores)
    tmax = max(times)
    value = cores * cells * args.steps / tmax
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tmax / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "grid": [H, W, L], "dt_s": dt, "members": members, "options": options or None},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample + "; numpy oracle pinned bit-exactly to the reference's outputs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_native(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gcmiipy_b200 import _lib, dynamics, geometry, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus)

    H, W, L, dt, members, desc = WORKLOADS[args.workload]
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    options = parse_options(args.options)
    if options:
        assert world == 1 or members == 1, "the opt-in terms run on whole grids and on latitude bands"
        dynamics.configure(geom, **options)     # before the BandStepper: its bands then carry 2 + 2 halo rows
        desc += " + opt-in terms"
    _lib.lib().gcm_pe25_select_path(args.path)
    for kv in args.knob:
        i, v = kv.split("=")
        _lib.lib().gcm_tuning_knob(int(i), int(v))
    if members > 1:
        per = members // world
        states = [synthetic.synthetic_state(geom, seed=1234 + rank * per + m) for m in range(min(per, 8))]
        reps = (per + len(states) - 1) // len(states)
        s0 = tuple(np.ascontiguousarray(np.concatenate([np.stack([st[f] for st in states])] * reps)[:per])
                   for f in range(5))
        cells_rank = per * H * W * L
        stepper = dynamics.Stepper(geom, *s0)
        scaling = "weak" if False else "strong"
    else:
        s0 = synthetic.synthetic_state(geom, seed=1234)
        if world > 1:
            from gcmiipy_b200 import bands
            stepper = bands.BandStepper(geom, *s0, rank=rank, world=world)
            cells_rank = stepper.owned_rows * W * L
        else:
            stepper = dynamics.Stepper(geom, *s0)
            cells_rank = H * W * L
        scaling = "strong"
    total_cells = H * W * L * members

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stepper_initial = [x.clone() for x in stepper.tensors()]

    def reset():
        stepper.upload(*stepper_initial)

    # ---- device-resident throughput ("value") -------------------------------------------------------
    stepper.step(dt, args.warmup)
    # two more untimed calls of the timed call's own shape: a multi-step call replays step pairs as a CUDA graph that
    # is captured on first use, once for each of the two buffer parities (one-time cost, not steady state)
    for _ in range(2):
        stepper.step(dt, args.steps)
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rep_ms = []
    repeats = max(args.repeats, 1)
    while len(rep_ms) < repeats or (sum(rep_ms) < 200.0 and len(rep_ms) < 200):
        reset()
        barrier()
        e0.record()
        stepper.step(dt, args.steps)
        e1.record()
        barrier()
        tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)        # the slowest rank defines the repeat
        rep_ms.append(float(tms.item()))
    ms = sorted(rep_ms)[len(rep_ms) // 2]                       # median repeat: the sustained figure
    ms_best = min(rep_ms)
    finite = all(bool(torch.isfinite(x).all()) for x in stepper.tensors())
    value = total_cells * args.steps / (ms * 1e-3)
    # SHA-256 of the whole state after K steps from the fixed initial state (gathered over the ranks): the same line
    # at every N proves the band decomposition bitwise invariant
    state_sha = None
    if not args.no_hash:
        import hashlib
        if world > 1:
            full = stepper.gather()
        else:
            full = stepper.download()
        if rank == 0:
            h = hashlib.sha256()
            for a in full:
                h.update(np.ascontiguousarray(a.cpu().numpy() if hasattr(a, "cpu") else a).tobytes())
            state_sha = h.hexdigest()
        del full

    # ---- end to end through the host-facing call: pinned host state in, pinned host state out, every step ----
    reset()
    host_in = [x.cpu().pin_memory() for x in stepper.tensors()]
    host_out = [torch.empty_like(x).pin_memory() for x in host_in]
    h2d = sum(x.numel() * 8 for x in host_in)
    pipelined = world == 1 and members == 1 and not args.e2e_serial     # steps overlap across calls (host_step.cu)
    kw = {"pipelined": True} if pipelined else {}
    for _ in range(2):
        stepper.step_host(host_in, host_out, dt, 1, **kw)
    if pipelined:
        stepper.host_join()
    barrier()
    e2e_ms = []
    for _ in range(3):
        barrier()
        e0.record()
        for _ in range(args.steps):
            stepper.step_host(host_in, host_out, dt, 1, **kw)
            host_in, host_out = host_out, host_in
        if pipelined:
            stepper.host_join()          # the last copy-outs are inside the timed region
        e1.record()
        barrier()
        tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        e2e_ms.append(float(tms.item()))
    e2e_value = total_cells * args.steps / (sorted(e2e_ms)[1] * 1e-3)            # median of three
    # clocks / throttle reasons sampled across both timed regions (the device-resident one alone lasts ~50 ms)
    clocks = sampler.stop() if sampler else None

    # ---- per-kernel timing pass (CUDA events on the launching stream) for the roofline of the dominant kernel ----
    roofline, launches_per_step = None, None
    lib = _lib.lib()
    reset()
    nk = lib.gcm_prof_kinds()
    import ctypes
    k_ms = (ctypes.c_double * nk)()
    k_n = (ctypes.c_longlong * nk)()
    psteps = min(args.steps, 10)
    stepper.step(dt, 2)
    torch.cuda.synchronize()
    lib.gcm_prof_enable(1)
    lib.gcm_prof_collect(k_ms, k_n)
    stepper.step(dt, psteps)
    lib.gcm_prof_collect(k_ms, k_n)
    lib.gcm_prof_enable(0)
    kinds = [(lib.gcm_prof_kind_name(k).decode(), k_ms[k], k_n[k]) for k in range(nk) if k_n[k] > 0]
    if kinds:
        launches_per_step = sum(n for _, _, n in kinds) / psteps
        name, tot, n = max((x for x in kinds if x[0] != "halo_exchange"), key=lambda x: x[1])
        peak, peak_src = measured_peak()
        launches_of_kind_per_step = n / psteps
        alg_bytes = b_alg(L) * cells_rank / launches_of_kind_per_step      # the kernel's share of one cell-update
        achieved = alg_bytes / (tot / n * 1e-3) / 1e9
        traffic = ncu_traffic(name) if (args.workload == "c5" and world == 1) else None
        halo = [(t, n) for nm, t, n in kinds if nm == "halo_exchange"]
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic["bytes_per_launch"] if traffic else None,
                    # `traffic` is NOT measured by this run: it is the ncu --set full capture named here (profiles/)
                    "traffic_source": traffic["source"] if traffic else None, "peak_source": peak_src,
                    "halo_exchange_ms_per_step": (halo[0][0] / psteps) if halo else None,
                    "avg_launch_ms": tot / n, "alg_bytes_per_launch": alg_bytes,
                    # the two chains of the row phase run side by side, so the per-kernel times add up to more than
                    # the step: the share is taken against the measured step, the second figure against that sum
                    "kernel_share_of_step": (tot / psteps) / (ms / args.steps),
                    "kernel_share_of_summed_kernel_time": tot / sum(t for _, t, _ in kinds),
                    "kernels_ms_per_step": {nm: t / psteps for nm, t, _ in kinds},
                    "whole_step": {"achieved": value * b_alg(L) / 1e9 / world, "frac": value * b_alg(L) / 1e9 / world / peak}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline: the oracle on this box's host cores, bounded sample --------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        # bounded sample of about 15 s of single-core work (the numpy port does ~2.3 M cell-updates/s)
        if members > 1:
            nmem = max(64, int(15 * 2.3e6 // (H * W * L)))
            t = cpu_oracle_step_rate(H, W, L, dt, nmem, options=options)
            cpu = {"value": H * W * L * nmem / t, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "%d member-steps of the %dx%dx%d grid, numpy oracle, 1 thread" % (nmem, W, H, L)}
        else:
            rows = H if H * W * L <= 2_000_000 else 180
            nst = max(1, int(15 * 2.3e6 // (rows * W * L)))
            t = cpu_oracle_step_rate(rows, W, L, dt, nst, options=options)
            cpu = {"value": rows * W * L * nst / t, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "%d Matsuno step(s) of a %dx%dx%d grid, numpy oracle, 1 thread" % (nst, W, rows, L)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "best_ms_per_step": ms_best / args.steps, "repeats": len(rep_ms),
        "timed_ms_total": sum(rep_ms), "value_is": "median of `repeats` timed calls of `steps` steps each",
        "state_sha256": state_sha, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "grid": [H, W, L], "dt_s": dt, "members": members, "options": options or None,
                   "parallelism": "lat-bands x%d" % world if members == 1 else "members split x%d" % world,
                   "l2_policy": "state %.0f MB > 126 MB L2: inputs larger than L2, no flush" % (total_cells * b_alg(L) / 2e6 / world)
                   if total_cells * b_alg(L) / 2 / world > 126e6 else "state fits L2 (latency/ALU-bound config); no flush"},
        "finite": finite,
        "halo_transport": (("peer mailboxes (NVLink peer stores, CUDA IPC)" if getattr(stepper, "peer", False) else "NCCL send/recv")
                           if world > 1 and members == 1 else None),
        "peer_timeouts": stepper.peer_timeouts() if world > 1 and members == 1 and getattr(stepper, "comm", None) else None,
        "sim_days_per_day": (args.steps / (ms * 1e-3)) * dt,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d,
                "call": "Stepper.step_host(pinned host state in, pinned host state out) every step"
                        + (", pipelined across steps, joined inside the timed region" if pipelined else "")},
        "gpu_launches": int(round((launches_per_step or 0) * args.steps)),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _cpu_2d(kind, H, W, dt, dx, nsteps, seed=7):
    """The oracle's 2-D scheme on one host core; returns (seconds, state)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import np_oracle as O
    s = _state_2d(kind, H, W, seed)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        s = O.matsumo_scheme(*s, dx, dt) if kind == "sw2d" else O.pe2d_matsuno_timestep(*s, dt, dx)
    return time.perf_counter() - t0, s


def _state_2d(kind, H, W, seed=7):
    import numpy as np
    rng = np.random.default_rng(seed)
    if kind == "sw2d":           # matsuno_c_grid.py:146-157: flat 8000 m layer, u bump, + 0.1 m/s noise
        u = 0.1 * rng.standard_normal((H, W)); v = 0.1 * rng.standard_normal((H, W))
        u[H // 2, W // 2] += 1.0
        return u, v, np.full((H, W), 8000.0)
    p = 1e5 + 50.0 * rng.standard_normal((H, W))      # no_limits_2d.py:134-150 style: p, u, v, t, q
    return (p, 0.1 * rng.standard_normal((H, W)), 0.1 * rng.standard_normal((H, W)),
            300.0 + 0.5 * rng.standard_normal((H, W)), np.full((H, W), 1e-3))


def run_2d(args):
    """c1 / c1big: matsuno_c_grid.matsumo_scheme (48 B per cell-update: u, v, h read + written once per step);
    p2d: no_limits_2d.matsuno_timestep (64 B: p, u, v, t; q untouched).  Single GPU."""
    kind, H, W, dt, dx, desc = WORKLOADS_2D[args.workload]
    if args.workload == "c1dt700":
        args.steps = min(args.steps, 8)          # the scheme diverges after 8 steps at this dt
    bpc = 48.0 if kind == "sw2d" else 64.0
    cells = H * W
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        rows = H if cells <= 1 << 20 else max(8, (1 << 20) // W)
        n = max(1, min(args.steps, int(20 * 2.5e6 // (rows * W)) or 1))
        t, _ = _cpu_2d(kind, rows, W, dt, dx, n)
        v = rows * W * n / t
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": n, "warmup": 0, "ms_per_step": 1e3 * t / n, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": desc, "grid": [H, W], "dt_s": dt},
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                           "sample": "%d step(s) of a %dx%d grid, numpy oracle, 1 thread" % (n, W, rows)},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    import numpy as np
    import torch
    from gcmiipy_b200 import _lib, matsuno_c_grid, no_limits_2d
    assert args.gpus == 1 and torch.cuda.is_available(), "the 2-D workloads are single-GPU"
    torch.cuda.set_device(0)
    s0 = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in _state_2d(kind, H, W)]
    step = (lambda s, n: matsuno_c_grid.matsumo_scheme(*s, dx, dt, nsteps=n)) if kind == "sw2d" else (
        lambda s, n: no_limits_2d.matsuno_timestep(*s, dt, dx, nsteps=n))
    for _ in range(3):
        step(s0, max(args.warmup, 1))
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    rep = []
    while len(rep) < max(args.repeats, 1) or (sum(rep) < 200.0 and len(rep) < 500):
        torch.cuda.synchronize()
        e0.record()
        out = step(s0, args.steps)
        e1.record()
        torch.cuda.synchronize()
        rep.append(e0.elapsed_time(e1))
    ms = sorted(rep)[len(rep) // 2]
    value = cells * args.steps / (ms * 1e-3)
    # e2e: pinned host arrays in, pinned host arrays out, EVERY step (the reference's own calling convention: numpy in,
    # numpy out): H2D of the step's inputs, the step, D2H of its result inside the timed region
    hin = [a.cpu().pin_memory() for a in s0]
    hout = [torch.empty_like(a).pin_memory() for a in hin]

    def host_step():
        dev = [h.cuda(non_blocking=True) for h in hin]
        res = step(dev, 1)
        for h, r in zip(hout, res):
            h.copy_(r, non_blocking=True)

    for _ in range(2):
        host_step()
    n_e2e = max(3, min(args.steps, 20))
    torch.cuda.synchronize()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    for _ in range(n_e2e):
        host_step()
        hin, hout = hout, hin
    ee1.record()
    torch.cuda.synchronize()
    e2e = cells * n_e2e / (ee0.elapsed_time(ee1) * 1e-3)
    clocks = sampler.stop()
    peak, peak_src = measured_peak()
    achieved = value * bpc / 1e9
    rows = H if cells <= 1 << 20 else max(8, (1 << 20) // W)
    ncpu = max(1, int(10 * 2.5e6 // (rows * W)))
    tcpu, _ = _cpu_2d(kind, rows, W, dt, dx, ncpu)
    nfields = 3 if kind == "sw2d" else 5
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "best_ms_per_step": min(rep) / args.steps, "repeats": len(rep),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "grid": [H, W], "dt_s": dt, "dx_m": dx,
                       "l2_policy": "state %.0f MB > 126 MB L2: no flush" % (cells * bpc / 2e6) if cells * bpc / 2 > 126e6
                       else "state fits one SM's shared memory / L2 (latency-bound config); no flush"},
            "finite": all(bool(torch.isfinite(x).all()) for x in out),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": cells * 8 * nfields,
                    "d2h_bytes_per_step": cells * 8 * nfields},
            "gpu_launches": args.steps if not (kind == "sw2d" and 6 * cells * 8 <= 220 * 1024) else 1,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "sw2d_matsuno_tile_kernel" if kind == "sw2d" else "pe2d_half_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": peak_src, "alg_bytes_per_cell_update": bpc},
            "cpu_baseline": {"value": rows * W * ncpu / tcpu, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": "%d step(s) of a %dx%d grid, numpy oracle, 1 thread" % (ncpu, W, rows)}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS) + sorted(WORKLOADS_2D))
    ap.add_argument("--repeats", type=int, default=5, help="timed calls of --steps steps (median reported; at least 200 ms)")
    ap.add_argument("--no-hash", action="store_true", help="skip the SHA-256 of the final state")
    ap.add_argument("--e2e-serial", action="store_true", help="e2e through the joining host step (no overlap across steps)")
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--path", type=int, default=0, help="0 = fused kernels (default), 1 = general 4-kernel path")
    ap.add_argument("--options", default="", help="opt-in terms of the step: coriolis,limit_q,limit_t,viscosity=NU "
                    "(BASELINE configs[1] names flux limiter + viscosity: --workload c2 --options limit_q,limit_t,viscosity=1e5)")
    ap.add_argument("--knob", action="append", default=[], help="tuning knob i=v (gcm_tuning_knob)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.workload in WORKLOADS_2D:
        run_2d(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
