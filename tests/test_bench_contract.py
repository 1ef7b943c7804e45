"""bench.py prints exactly one JSON line on stdout with the keys of the benchmark contract (reference arm on CPU here,
our arm on a GPU)."""
import json
import os
import subprocess
import sys

import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must carry one JSON line, got {len(lines)}"
    return json.loads(lines[0])


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref (the unmodified reference) was not built")
def test_reference_arm_line():
    d = _run(["--impl", "reference", "--n-per-dim", "14", "--steps", "10", "--warmup", "1"])
    assert d["impl"] == "reference" and COMMON <= set(d)
    assert d["unit"] == "MFUPs/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_own_arm_line():
    d = _run(["--n-per-dim", "40", "--steps", "10", "--warmup", "3", "--e2e-steps", "10", "--no-cpu-baseline"])
    assert COMMON | {"roofline", "gpu_launches", "clocks"} <= set(d)
    assert d["unit"] == "MFUPs/s" and d["value"] > 0 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["steps"] == 10 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 1
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 3 * 8 * 40 ** 3 and e["d2h_bytes_per_step"] >= 3 * 8 * 40 ** 3
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
