"""Synthetic initial states for benchmarks and smoke runs: the reference's own
initial conditions (no_limits_2_5d.py:146-168, :222-226) plus a smooth seeded band-limited perturbation
so that no term of the step is identically zero.  Host side, numpy."""
import numpy as np

from . import no_limits_2_5d


def synthetic_state(geom, seed=1234, amp_u=1.0, amp_p=50.0, amp_t=0.5):
    p, u, v, t, q, _ = no_limits_2_5d.gen_initial_conditions(geom)
    v[0, 0, 0] = 0.1
    u *= 0
    L, H, W = u.shape
    rng = np.random.default_rng(seed)

    def smooth(shape):
        f = np.fft.rfft2(rng.standard_normal(shape))
        nj = max(1, shape[-2] // 8)
        ni = max(1, shape[-1] // 8)
        f[..., nj + 1:shape[-2] - nj, :] = 0
        f[..., :, ni + 1:] = 0
        x = np.fft.irfft2(f, s=shape[-2:])
        m = np.max(np.abs(x))
        return x / m if m > 0 else x

    u = u + amp_u * smooth((L, H, W))
    v = v + amp_u * smooth((L, H, W))
    p = p + amp_p * smooth((H, W))
    t = t + amp_t * smooth((L, H, W))
    v[:, -1, :] = 0
    return p, u, v, t, q
