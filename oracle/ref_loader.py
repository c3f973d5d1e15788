"""Import the UNMODIFIED reference (/root/reference) through test-side stand-ins.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py and by tests that are
skipped when /root/reference is absent, i.e. on the GPU box).

Three stand-ins, zero edits to the reference:
  1. oracle/shims/pint        - SI-magnitude Quantity (pint is not installable here)
  2. oracle/shims/matplotlib  - no-op pyplot
  3. np.float = float         - removed from numpy >= 1.24, used e.g. matsuno_c_grid.py:153
"""
import contextlib
import importlib
import io
import os
import sys

import numpy as np

REFERENCE_DIR = os.environ.get("GCMIIPY_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "dynamics.py"))


def load(*names):
    """Return the named reference modules (e.g. load('dynamics', 'geometry'))."""
    if not available():
        raise RuntimeError("reference not present at %s" % REFERENCE_DIR)
    if not hasattr(np, "float"):
        np.float = float  # noqa: NPY001 - restores the alias the reference relies on
    for p in (REFERENCE_DIR, _SHIMS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_DIR)
    sys.path.insert(0, _SHIMS)
    mods = []
    with quiet():
        for n in names:
            mods.append(importlib.import_module(n))
    return mods[0] if len(mods) == 1 else tuple(mods)


@contextlib.contextmanager
def quiet():
    """The reference prints inside the hot path (dynamics.py:137-138, geometry.py:117-122)."""
    with contextlib.redirect_stdout(io.StringIO()):
        yield
