"""bench.py prints exactly one JSON line with the keys the driver reads -- checked on a reduced particle count
(--n) so that it runs in seconds: the reference arm on the CPU (not gpu), our arm on the GPU (gpu)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines          # ONE line on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--n", "6000", "--steps", "2", "--warmup", "1", "--cpu-seconds", "0.3")
    assert BASE <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "biot_savart_pair_interactions_per_sec" and d["unit"] == "interactions/s" and d["higher_is_better"] is True
    assert d["value"] > 1e6 and d["steps"] == 2 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_our_arm_line():
    d = run_bench("--n", "40000", "--steps", "3", "--warmup", "3", "--cpu-seconds", "0.5", "--no-variants")
    assert BASE | {"roofline", "clocks"} <= set(d) and "impl" not in d
    assert d["metric"] == "biot_savart_pair_interactions_per_sec" and d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3
    assert d["value"] > 1e11 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0.0 < r["frac"] < 1.5 and r["unit"] == "TFLOP/s"
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["max_rel_err_gpu_vs_cpu"] < 1e-5
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["config"]["results"] == 2 and "velocity + gradient" in d["config"]["workload"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
