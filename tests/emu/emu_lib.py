"""TEST INFRASTRUCTURE ONLY: load the kernel sources compiled for the CPU emulator (tests/emu/build_emu.py)."""
import ctypes
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_emu  # noqa: E402

_CDLL = None


def load():
    global _CDLL
    if _CDLL is None:
        _CDLL = ctypes.CDLL(build_emu.build())
    return _CDLL
