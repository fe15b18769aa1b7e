"""GPU test (-m gpu) of the native bench.py arm on the tiny self-test shape: one JSON line with the contract's keys,
kernels actually launched, e2e measured through the public API with host buffers."""

import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_native_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--config", "tiny", "--steps", "3", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["metric"] == "fold matrices/sec" and d["unit"] == "fold-matrices/s" and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["steps"] == 3 and d["warmup"] == 3
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    roof = d["roofline"]
    assert roof["bound"] in ("hbm", "tensor") and roof["peak"] > 0 and roof["achieved"] > 0 and roof["kernel"].startswith("k_gram")
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    par = d["parity"]
    assert par["stats_bit_exact"] is True and par["xtx"] <= 1e-12 and par["xty"] <= 1e-12 and par["joint"] <= 1e-12 and par["folds_checked"] == 4
    assert d["parity_e2e_path"]["stats_bit_exact"] is True and d["parity_e2e_path"]["joint"] <= 1e-12
