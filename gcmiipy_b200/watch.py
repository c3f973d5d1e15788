"""Asynchronous error reporting of the C-ABI layer ("error conventions").

The reference's callers poll `np.isnan(u).any()` between steps (matsuno_c_grid.py:184-187) and print min / max
(no_limits_2_5d.py:85-88) -- a device -> host round trip and a full-array reduction per step.  Here every step kernel
bumps a 32-bit counter on the device when it writes an inf or NaN; `NonFiniteWatch.poll()` enqueues a 4-byte copy of
it into pinned host memory behind the steps already queued and returns at once, `ready()` / `value()` look at the copy
without synchronising the stream.
"""
import ctypes

import torch

from . import _lib

__all__ = ["NonFiniteWatch", "last_status"]


def last_status(clear=False):
    """Newest non-zero status any entry point returned on this thread (0 = none), see gcm_last_status."""
    return int(_lib.lib().gcm_last_status(1 if clear else 0))


class NonFiniteWatch:
    """
    >>> watch = NonFiniteWatch()          # resets the device's counter
    >>> stepper.step(dt, 100); watch.poll()
    >>> ...                               # more work, no synchronisation
    >>> if watch.ready() and watch.value(): raise FloatingPointError("model blew up")
    """

    def __init__(self, reset=True):
        cuda = _lib.device().type == "cuda"
        self._host = torch.zeros(1, dtype=torch.int32)
        if cuda:
            self._host = self._host.pin_memory()
        self._event = torch.cuda.Event() if cuda else None
        self._polled = False
        if reset:
            _lib.check(_lib.lib().gcm_nonfinite_read(None, 1, _lib.stream()), "gcm_nonfinite_read")

    def poll(self, reset=False):
        """Enqueue a read of the counter on the current stream (and optionally zero it afterwards)."""
        _lib.check(_lib.lib().gcm_nonfinite_read(ctypes.c_void_p(self._host.data_ptr()), 1 if reset else 0,
                                                 _lib.stream()), "gcm_nonfinite_read")
        if self._event is not None:
            self._event.record()
        self._polled = True

    def ready(self):
        """True once the last poll() has landed in host memory (never blocks)."""
        return self._polled and (self._event is None or self._event.query())

    def value(self, wait=False):
        """Threads that wrote a non-finite value up to the last poll(); None if that poll has not landed yet
        (wait=True blocks on it instead)."""
        if not self._polled:
            return None
        if self._event is not None:
            if wait:
                self._event.synchronize()
            elif not self._event.query():
                return None
        return int(self._host.item()) & 0xFFFFFFFF
