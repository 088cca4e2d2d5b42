"""CPU tests of bench.py's contract pieces that need no GPU: the reference arm's JSON line for every BASELINE config, and
that both arms describe the workload with the SAME `config` object (the driver's `same_config`)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.mark.parametrize("cfg", [2, 3, 4, 5])
def test_reference_arm_line(cfg, oracle):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", str(cfg), "--batch", "16",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # exactly ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert set(d["cpu_baseline"]["reference_solver_probe"]) == {"casadi", "reference_install", "reference_checkout"}
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["dtype"] == "f64" and d["config"]["baseline_config"] == cfg
    import bench
    w = bench.WORKS[cfg](16, 1)
    ours = w.config(16, 1, w.default_inflight, bench.exchange_name(1, None))
    assert ours == d["config"]                          # the product arm builds its `config` with the same call
    assert d["metric"] == w.metric and w.alg_bytes > 0


def test_default_invocation_is_the_metric_config():
    import bench
    assert bench.WORKS[2].metric.startswith("MPC-CBF solves/sec (N=20, 6-state, 3 obs)")
    w = bench.WORKS[2](4, 1)
    assert w.rec.shape == (4, 142) and w.alg_bytes == 1136 + 536
