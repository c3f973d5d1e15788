"""Marshalling between the reference's array families and device memory.

The reference passes pint Quantities wrapping float64 numpy arrays (constants.py:2-5).  This layer
accepts pint Quantities (converted with .to_base_units()), plain numpy arrays / Python floats
(assumed SI) or torch tensors already on the device, and hands results back in the same family.
Layout contract: float64, C-contiguous, 2-D [j, i], 3-D [k, j, i] (i fastest).
"""
import ctypes

import numpy as np
import torch

from . import _lib


def is_quantity(x):
    return hasattr(x, "to_base_units") and hasattr(x, "magnitude")


def magnitude(x):
    """SI-base magnitude of a Quantity, or x itself."""
    if is_quantity(x):
        return np.asarray(x.to_base_units().magnitude, dtype=np.float64)
    return x


def scalar(x):
    """Python float (SI) of a Quantity / 0-d array / number."""
    m = magnitude(x)
    if isinstance(m, torch.Tensor):
        return float(m.item())
    return float(np.asarray(m, dtype=np.float64).reshape(-1)[0]) if np.ndim(m) else float(m)


class Family:
    """Remembers which array family the caller used, to hand results back in it."""

    def __init__(self, *inputs):
        self.quantity = next((x for x in inputs if is_quantity(x)), None)
        self.torch = any(isinstance(x, torch.Tensor) for x in inputs)

    def out(self, t, unit=None):
        if self.torch and self.quantity is None:
            return t
        a = t.detach().cpu().numpy()
        if self.quantity is not None:
            return wrap_quantity(self.quantity, a, unit)
        return a


def wrap_quantity(like, a, unit):
    reg = getattr(like, "_REGISTRY", None)
    if reg is not None and unit is not None:       # real pint
        return reg.Quantity(a, unit)
    try:
        return type(like)(a)                       # SI stand-in Quantity
    except Exception:
        return a


def dev(x, shape=None):
    """float64 contiguous tensor on the device holding x (copy only when needed)."""
    d = _lib.device()
    if isinstance(x, torch.Tensor):
        t = x.to(device=d, dtype=torch.float64)
    else:
        a = np.require(np.asarray(magnitude(x), dtype=np.float64), requirements=["C", "W", "A"])  # keeps 0-d
        t = torch.from_numpy(a).to(d)
    if shape is not None and tuple(t.shape) != tuple(shape):
        t = t.expand(shape) if t.dim() == len(shape) else t.reshape(shape)
    return t.contiguous()


def empty(shape):
    return torch.empty(tuple(shape), dtype=torch.float64, device=_lib.device())


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def hptr(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None
