#!/usr/bin/env python
"""Benchmark of the Matsuno C-grid hot path (BASELINE.json: cell-updates/s, fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5|c3|c2|c4] [--impl native|reference]
                    [--options coriolis,limit_q,limit_t,viscosity=NU]

One "step" = one full Matsuno step (predictor + corrector, dynamics.py:230-237) of the 2.5-D model over the
whole grid.  Default workload: the 0.25 deg grid 1440 x 720 x 9 (BASELINE.json configs[4], the grid the metric
"cell-updates/s at 1/2/4/8 B200" is quoted on; 307 MB of state > 126 MB L2, so every step streams from HBM).
N > 1 (torchrun, one rank per GPU): latitude-band decomposition, strong scaling.
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
    "c5b8": (90, 1440, 9, 10.0, 1, "1440x90x9: one of 8 latitude bands of configs[4], stepped alone (tuning aid)"),
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
    reset()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    stepper.step(dt, args.steps)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    finite = all(bool(torch.isfinite(x).all()) for x in stepper.tensors())
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    value = total_cells * args.steps / (ms * 1e-3)

    # ---- end to end through the host-facing call: pinned host state in, pinned host state out, every step ----
    reset()
    host_in = [x.cpu().pin_memory() for x in stepper.tensors()]
    host_out = [torch.empty_like(x).pin_memory() for x in host_in]
    h2d = sum(x.numel() * 8 for x in host_in)
    for _ in range(2):
        stepper.step_host(host_in, host_out, dt, 1)
    barrier()
    e0.record()
    for _ in range(args.steps):
        stepper.step_host(host_in, host_out, dt, 1)
        host_in, host_out = host_out, host_in
    e1.record()
    barrier()
    tms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_value = total_cells * args.steps / (float(tms.item()) * 1e-3)
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
        name, tot, n = max(kinds, key=lambda x: x[1])
        peak, peak_src = measured_peak()
        launches_of_kind_per_step = n / psteps
        alg_bytes = b_alg(L) * cells_rank / launches_of_kind_per_step      # the kernel's share of one cell-update
        achieved = alg_bytes / (tot / n * 1e-3) / 1e9
        traffic = ncu_traffic(name) if (args.workload == "c5" and world == 1) else None
        roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic["bytes_per_launch"] if traffic else None,
                    "traffic_source": traffic["source"] if traffic else None, "peak_source": peak_src,
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
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "grid": [H, W, L], "dt_s": dt, "members": members, "options": options or None,
                   "parallelism": "lat-bands x%d" % world if members == 1 else "members split x%d" % world,
                   "l2_policy": "state %.0f MB > 126 MB L2: inputs larger than L2, no flush" % (total_cells * b_alg(L) / 2e6 / world)
                   if total_cells * b_alg(L) / 2 / world > 126e6 else "state fits L2 (latency/ALU-bound config); no flush"},
        "finite": finite,
        "sim_days_per_day": (args.steps / (ms * 1e-3)) * dt,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d},
        "gpu_launches": int(round((launches_per_step or 0) * args.steps)),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--path", type=int, default=0, help="0 = fused kernels (default), 1 = general 4-kernel path")
    ap.add_argument("--options", default="", help="opt-in terms of the step: coriolis,limit_q,limit_t,viscosity=NU "
                    "(BASELINE configs[1] names flux limiter + viscosity: --workload c2 --options limit_q,limit_t,viscosity=1e5)")
    ap.add_argument("--knob", action="append", default=[], help="tuning knob i=v (gcm_tuning_knob)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
