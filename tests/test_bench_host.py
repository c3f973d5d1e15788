"""Host-side logic of bench.py that needs no GPU: option parsing, the algorithmic-bytes figure of SURVEY.md section 8d,
the lookup of the committed ncu traffic, and the reference arm (the pinned numpy oracle on the host cores)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_per_cell_update():
    # 4 three-D fields in + 4 out + the 2-D p in + out, fp64: 64 + 16 / L
    assert bench.b_alg(9) == pytest.approx(65.7777777, rel=1e-6)
    assert bench.b_alg(1) == 80.0


def test_parse_options():
    assert bench.parse_options("") == {}
    assert bench.parse_options("coriolis,limit_q,viscosity=1e5") == {"coriolis": True, "limit_q": True, "viscosity": 1e5}
    with pytest.raises(SystemExit):
        bench.parse_options("flux_capacitor")


def test_every_workload_names_a_baseline_config():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert len(base["configs"]) == 5
    for key, idx in (("c2", 1), ("c3", 2), ("c4", 3), ("c5", 4)):
        assert "configs[%d]" % idx in bench.WORKLOADS[key][5]
    H, W, L = bench.WORKLOADS["c5"][:3]
    assert (W, H, L) == (1440, 720, 9)


def test_ncu_traffic_comes_from_the_newest_committed_summary():
    t = bench.ncu_traffic("pe25f_update_tiled_kernel")
    assert t is not None and t["source"].startswith("profiles/") and t["bytes_per_launch"] > 3.0e8


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cell_updates_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0


def test_two_d_workloads_and_states():
    """c1 / c1dt700 / c1big / p2d: BASELINE configs[0] and the 2-D primitive-equation scheme; seeded states."""
    import numpy as np
    assert bench.WORKLOADS_2D["c1"][:3] == ("sw2d", 64, 64) and bench.WORKLOADS_2D["c1dt700"][3] == 700.0
    assert "configs[0]" in bench.WORKLOADS_2D["c1"][5]
    u, v, h = bench._state_2d("sw2d", 8, 8)
    assert h.shape == (8, 8) and np.all(h == 8000.0) and abs(u[4, 4]) > 0.5
    s = bench._state_2d("pe2d", 6, 10)
    assert len(s) == 5 and all(a.shape == (6, 10) for a in s)
    t, out = bench._cpu_2d("sw2d", 8, 8, 300.0, 300e3, 2)
    assert t > 0 and all(np.isfinite(a).all() for a in out)
