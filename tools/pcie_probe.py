#!/usr/bin/env python
"""PCIe ceiling of the box: pinned host <-> device copies of the size bench.py's e2e moves per step (307 MB), one
direction at a time and both at once on two streams.  The e2e figure of bench.py is bounded by the last number."""
import torch

n = 306892800 // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s1, s2):
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


gb = n * 8 / 1e9
for name, fn in (("H2D", h2d), ("D2H", d2h), ("both", both)):
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    ms = timed(fn)
    print("%s: %.2f ms per 307 MB  = %.1f GB/s per direction" % (name, ms, gb / (ms * 1e-3)))
