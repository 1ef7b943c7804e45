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


def test_mt19937_uniform_matches_libstdcxx():
    """bench.py's C2 jitter = std::uniform_real_distribution<double>(-0.15, 0.15) on std::mt19937(42);
    the literals are the first draws of that C++ program (g++ 13, libstdc++)."""
    sys.path.insert(0, ROOT)
    import bench
    got = bench.mt19937_uniform(42, 4, -0.15, 0.15)
    want = [0.088962895286353788, -0.094969563631989454, 0.083907299283798364, 0.029055048474016965]
    assert list(got) == want, list(got)


@pytest.mark.skipif(not oracle.have_refbench(), reason="oracle/_ref (the unmodified reference) was not built")
@pytest.mark.parametrize("workload", ["c2", "c3"])
def test_reference_arm_line(workload):
    d = _run(["--impl", "reference", "--workload", workload, "--n-per-dim", "14", "--steps", "10", "--warmup", "1"])
    assert d["impl"] == "reference" and COMMON <= set(d)
    assert d["config"]["host_threads"] == len(os.sched_getaffinity(0))
    assert {q["functor"] for q in d["config"]["configurations_timed"]} == {"LJFunctor", "LJFunctorHWY"}
    assert d["unit"] == "MFUPs/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


@pytest.mark.gpu
def test_own_arm_line():
    d = _run(["--workload", "c2", "--n-per-dim", "40", "--steps", "10", "--warmup", "3", "--e2e-steps", "10",
              "--no-cpu-baseline"])
    assert COMMON | {"roofline", "gpu_launches", "clocks"} <= set(d)
    assert d["unit"] == "MFUPs/s" and d["value"] > 0 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["steps"] == 10 and d["warmup"] == 3 and d["gpu_launches"] > 0 and d["scaling"] == "weak"
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 1
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 3 * 8 * 40 ** 3 and e["d2h_bytes_per_step"] >= 3 * 8 * 40 ** 3
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])


@pytest.mark.gpu
def test_default_workload_is_c3_with_c2_alongside():
    """The default line is the north_star configuration (C3, strong scaling) and carries the C2 measurement as well."""
    d = _run(["--n-per-dim", "64", "--steps", "20", "--warmup", "5", "--e2e-steps", "10", "--no-cpu-baseline"])
    assert d["scaling"] == "strong" and d["steps"] == 20 and d["warmup"] == 5
    assert d["config"]["workload"].startswith("C3") and d["config"]["rebuilds_in_timed_region"] == 2
    assert "phases_ms_per_step" in d and d["roofline"]["frac"] > 0
    c2 = d["c2"]
    assert c2["config"]["workload"].startswith("C2") and c2["value"] > 0 and c2["e2e"]["value"] > 0
    assert c2["config"]["particles_total"] == 1000000
