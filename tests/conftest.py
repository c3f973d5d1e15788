import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


# The parity cases run against two builds of the SAME kernel sources:
#   "gpu": gcmiipy_b200/_lib/libgcm_b200.so on a real B200 (-m gpu) -- the parity tests proper;
#   "emu": the sources compiled for the CPU emulator in tests/emu (no GPU in the build container): checks
#          indexing / halo / scan logic and the Python host layer before any GPU minute is spent.
@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    import torch
    from gcmiipy_b200 import _lib
    if request.param == "emu":
        from emu import emu_lib
        _lib._override_for_tests(emu_lib.load(), torch.device("cpu"))
        yield "emu"
        _lib._override_for_tests(None, None)
    else:
        assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
        _lib._override_for_tests(None, None)
        assert _lib.lib() is not None
        yield "gpu"
