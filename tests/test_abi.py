"""The C-ABI library builds, loads and exports every symbol include/gcm_b200.h declares (no compute calls:
this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gcm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gcm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("gcm_geom_create", "gcm_pe25_half_step", "gcm_pe25_matsuno_step", "gcm_sw2d_matsuno_step",
                 "gcm_pe2d_matsuno_step", "gcm_polar_filter", "gcm_phi_port_pgf", "gcm_laplacian5",
                 "gcm_fl_donor_cell_advection", "gcm_halo_pack"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import gcmiipy_b200
    from gcmiipy_b200 import _abi
    lib = gcmiipy_b200.build()
    cdll = ctypes.CDLL(lib)
    syms = declared_symbols()
    missing = [s for s in syms if not hasattr(cdll, s)]
    assert not missing, missing
    assert sorted(_abi.SIGNATURES) == syms, set(_abi.SIGNATURES) ^ set(syms)
    _abi.bind(cdll)
    assert cdll.gcm_version() >= 100
    assert cdll.gcm_status_string(-2) == b"bad extent or row range"


def test_product_has_no_cpu_fallback():
    """Without a CUDA device the product refuses to compute; it never imports the oracle."""
    import torch
    from gcmiipy_b200 import _lib
    _lib._override_for_tests(None, None)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            _lib.device()
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gcmiipy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "np_oracle" not in text and "import oracle" not in text, f


def test_argument_errors_are_status_codes():
    """Bad arguments come back as negative status codes before any CUDA call (error convention of the ABI)."""
    import gcmiipy_b200
    from gcmiipy_b200 import _abi
    cdll = _abi.bind(ctypes.CDLL(gcmiipy_b200.build()))
    assert cdll.gcm_geom_create(None, None) == -1
    assert cdll.gcm_sw2d_operator(9, None, None, None, ctypes.c_void_p(16), 4, 4, 1.0, None) == -4
    assert cdll.gcm_sw2d_matsuno_step(None, None, None, None, None, None, 4, 4, 1.0, 1.0, 1, None, 0, None) == -1
    assert cdll.gcm_fl_calc_r(ctypes.c_void_p(16), ctypes.c_void_p(16), 0, 4, None) == -2
    assert cdll.gcm_pe25_workspace_bytes(None, 1) == 0
    desc = _abi.GeomDesc(H=4, W=3, L=2)
    h = ctypes.c_void_p()
    assert cdll.gcm_geom_create(ctypes.byref(desc), ctypes.byref(h)) == -1     # tables missing
    # entry points added for the band / host-resident paths: NULL handles are argument errors, nothing is launched
    assert cdll.gcm_pe25_half_step_rows(None, None, None, None, 1.0, 1, None, 0, None, None, None) == -1
    assert cdll.gcm_band_matsuno_step(None, None, None, None, None, 1.0, 1, 0, None, 0, None) == -1
    assert cdll.gcm_pe25_matsuno_step_host(None, None, None, None, None, None, 1.0, 0, None, 0, None) == -1
    assert cdll.gcm_comm_create(0, 0, None, ctypes.byref(h)) == -2
    assert cdll.gcm_comm_unique_id(None) == -1
    assert cdll.gcm_tuning_knob(99, 1) == -2


def test_band_entry_points_reject_bad_geometry(backend):
    """gcm_band_matsuno_step wants a band geometry (wrap_j = 0), gcm_pe25_matsuno_step_host a whole grid, and
    gcm_pe25_half_step_rows row segments inside the stored rows."""
    import numpy as np
    import torch
    from gcmiipy_b200 import _host, _lib, bands, dynamics, geometry
    from gcmiipy_b200.dynamics import _struct, _workspace
    geom = geometry.gen_geometry(12, 16, 3, sig_func=geometry.manabe_sig)
    s = [np.ones((12, 16))] + [np.ones((3, 12, 16)) for _ in range(4)]
    st = dynamics.Stepper(geom, *s)
    ws, need = _workspace(st.dg, 1)
    sc, sn = _struct(st.cur), _struct(st.nxt)
    lib = _lib.lib()
    seg = lambda *v: (ctypes.c_int * 4)(*v)
    rc = lib.gcm_pe25_half_step_rows(st.dg.handle, ctypes.byref(sc), ctypes.byref(sc), ctypes.byref(sn), 1.0, 1,
                                     _host.ptr(ws), need, seg(0, 13, 0, 0), seg(0, 12, 0, 0), _lib.stream())
    assert rc == -2                                                      # 13 rows from row 0 of a 12-row grid
    band = bands.BandStepper(geom, *s, rank=0, world=1, native=True)
    rc = lib.gcm_band_matsuno_step(st.dg.handle, band.comm, ctypes.byref(sc), ctypes.byref(sc), ctypes.byref(sn), 1.0, 1,
                                   0, _host.ptr(ws), need, _lib.stream())
    assert rc == -4                                                      # whole-grid geometry: not a band
    bc, bn = _struct(band.cur), _struct(band.nxt)
    host = [torch.zeros_like(x, device="cpu") for x in band.cur]
    hs = _struct(host)
    wsb, needb = _workspace(band.dg, 1)
    rc = lib.gcm_pe25_matsuno_step_host(band.dg.handle, ctypes.byref(hs), ctypes.byref(hs), ctypes.byref(bc),
                                        ctypes.byref(bc), ctypes.byref(bn), 1.0, 0, _host.ptr(wsb), needb, _lib.stream())
    assert rc == -4                                                      # band geometry: not a whole grid


def test_sass_shows_tma_and_mbarrier_instructions():
    """The built sm_100a library carries the Blackwell-native load path: cp.async.bulk.tensor -> UTMALDG (update kernel on
    TMA box loads), cp.async.bulk -> UBLKCP (pipelined filter), mbarrier -> SYNCS.*, and the LDGSTS (cp.async) path of the
    default update kernel.  Needs cuobjdump (CUDA toolkit) and the built library; no GPU."""
    import shutil
    import subprocess
    from gcmiipy_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not here")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass or "SM100" in sass.upper() or "sm_100" in sass
    for mnemonic in ("UTMALDG.3D", "UBLKCP", "SYNCS.ARRIVE.TRANS64", "SYNCS.PHASECHK.TRANS64.TRYWAIT", "LDGSTS"):
        assert mnemonic in sass, mnemonic
