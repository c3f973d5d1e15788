"""Loader of the C-ABI library (hand-written sm_100a kernels).  There is NO CPU fallback: if the
library has not been built, or no CUDA device is present, every compute entry point raises."""
import ctypes
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libgcm_b200.so")

_LIB = None
_DEVICE = None


class GcmError(RuntimeError):
    """A C-ABI call returned a non-zero status (include/gcm_b200.h: < 0 argument error, > 0 cudaError_t)."""

    def __init__(self, status, what):
        self.status = status
        try:
            msg = lib().gcm_status_string(status).decode()
        except Exception:  # pragma: no cover
            msg = "?"
        super().__init__("%s failed: status %d (%s)" % (what, status, msg))


def lib():
    """The bound CDLL; built in-tree by `python -m gcmiipy_b200._build` (or __graft_entry__.build())."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("gcmiipy_b200: %s is missing -- run `python -m gcmiipy_b200._build` "
                               "(there is no CPU fallback)" % LIB_PATH)
        _LIB = _abi.bind(ctypes.CDLL(LIB_PATH))
    return _LIB


def device():
    """The torch device the state lives on: the current CUDA device.  Raises without one."""
    if _DEVICE is not None:
        return _DEVICE
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("gcmiipy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream():
    """cudaStream_t of torch's current stream (all work is enqueued there; no hidden syncs)."""
    if _DEVICE is not None and _DEVICE.type != "cuda":
        return None
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(status, what):
    if status != 0:
        raise GcmError(status, what)


def _override_for_tests(cdll, dev):
    """tests/ only: route the host layer to the kernel sources compiled for the CPU emulator
    (tests/emu) so that the no-GPU suite can exercise the host logic.  Never called by the product."""
    global _LIB, _DEVICE
    _LIB = _abi.bind(cdll) if cdll is not None else None
    _DEVICE = dev
