#!/usr/bin/env python
"""Multi-GPU check of the latitude-band stepper (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [--grid H W L] [--steps K] [--time]

1. R-rank band run == single-GPU whole-grid run, BIT for BIT (every rank also steps the whole grid on its own GPU),
   with and without the comm/compute overlap, native NCCL ring (csrc/comm.cu) and torch.distributed P2P path.
2. --time: ms per Matsuno step of the 0.25 deg grid (1440 x 720 x 9) per variant, device-timed, max over ranks.
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcmiipy_b200 import bands, dynamics, geometry, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs=3, default=None)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--time", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    H, W, L = args.grid or (24 * world, 96, 9)
    geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
    geom.heightmap = 100.0 * np.random.default_rng(7).random((H, W))
    s0 = synthetic.synthetic_state(geom, seed=1234)
    dt = 60.0 * 180.0 / H if H >= 180 else 450.0 * 24.0 / H
    whole = dynamics.Stepper(geom, *s0)
    whole.step(dt, args.steps)
    ref = whole.download()
    ok = True
    # (native loop, overlap: 0 none / 1 exchange under the corrector's interior update / 2 under the predictor's interior,
    #  one exchange per step, peer mailboxes)
    variants = ((True, 1, True, True), (True, 2, True, True), (True, 0, True, True), (True, 1, True, False),
                (True, 1, False, True), (True, 1, False, False), (True, 0, False, True), (False, 0, False, False))
    for native, overlap, wide, peer in variants:
        os.environ["GCM_BAND_PEER"] = "1" if peer else "0"      # peer mailboxes over NVLink, or the NCCL ring
        b = bands.BandStepper(geom, *s0, native=native, wide_halo=wide)
        b.overlap = overlap
        b.step(dt, args.steps)
        b.step(dt, 2)                                            # a second call: message numbers carry over
        got = b.gather()
        whole2 = dynamics.Stepper(geom, *s0)
        whole2.step(dt, args.steps + 2)
        ref = whole2.download()
        same = all(np.array_equal(a, r) for a, r in zip(got, ref))
        finite = all(np.isfinite(a).all() for a in got)
        tmo = b.peer_timeouts() if b.comm is not None else 0
        flag = torch.tensor([int(same and finite and tmo == 0)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("bitwise %s  native=%s overlap=%s one_exchange=%s transport=%s timeouts=%d  grid %dx%dx%d  ranks %d  steps %d" %
                  ("OK" if flag.item() else "MISMATCH", native, overlap, wide, "peer" if b.peer else "nccl", tmo, W, H, L,
                   world, args.steps + 2), flush=True)
        ok = ok and bool(flag.item())
        del b
    if args.time:
        H, W, L, dt = 720, 1440, 9, 10.0
        geom = geometry.gen_geometry(H, W, L, sig_func=geometry.manabe_sig)
        s0 = synthetic.synthetic_state(geom, seed=1234)
        for native, overlap, wide, peer in variants:
            os.environ["GCM_BAND_PEER"] = "1" if peer else "0"
            b = bands.BandStepper(geom, *s0, native=native, wide_halo=wide)
            b.overlap = overlap
            b.step(dt, 5)
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b.step(dt, 30)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / 30], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                print("time native=%s overlap=%s one_exchange=%s transport=%s: %.4f ms/step  (%.3e cell-updates/s on %d GPUs)" %
                      (native, overlap, wide, "peer" if b.peer else "nccl", t.item(), H * W * L / (t.item() * 1e-3), world),
                      flush=True)
            del b
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
