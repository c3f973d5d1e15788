"""gcmiipy_b200 -- B200-native (sm_100a) Matsuno C-grid dynamical-core time-stepper.

Drop-in for ONE path of marthinwurer/gcmiipy: the Matsuno forward-backward step on the Arakawa C-grid
and the operators it is built from.  The sub-modules carry the reference's own module and function
names (`dynamics`, `no_limits_2_5d`, `geometry`, `matsuno_c_grid`, `no_limits_2d`, `low_pass`,
`phi_port`, `viscosity`, `flux_limiter`, `temperature`, `coordinates*`, `constants`); underneath,
Python/PyTorch host code calls the C ABI of include/gcm_b200.h (hand-written CUDA kernels for sm_100a).
There is no CPU fallback: importing is free, computing without the built library or without a CUDA
device raises.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into gcmiipy_b200/_lib/libgcm_b200.so (in-tree)."""
    from ._build import build as _build
    return _build(force=force, verbose=verbose)
