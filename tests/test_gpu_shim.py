"""The C++ drop-in shim (autopas_b200/shim/GpuContainers.h) behind the UNMODIFIED AutoPas container / traversal / functor
interfaces: oracle/_ref/shim_test is compiled here against /root/reference (oracle/Makefile, target `shimtest`) and
travels to the GPU box as a binary. It runs the reference's LinkedCells + lc_c08 + LJFunctor on the host and the GPU
containers through the same virtual interface, and compares forces read through the container iterators, Upot and
virial (1e-12), region iterators, leavers of updateContainer and deleteParticle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "shim_test")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/shim_test was not built (reference tree absent)")
def test_cpp_shim_matches_reference_through_autopas_interfaces():
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    print(r.stderr)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "SHIM TEST PASSED" in r.stdout
    assert "vtk record:" in r.stdout and "identical" in r.stdout  # device-side checkpoint == the reference writer's statements


BIN_MS = os.path.join(ROOT, "oracle", "_ref", "shim_test_ms")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(BIN_MS), reason="oracle/_ref/shim_test_ms was not built (reference tree absent)")
def test_cpp_shim_multisite_functor_matches_reference():
    """GpuLJMultisiteFunctor around mdLib::LJMultisiteFunctor in the reference's MULTISITE build mode: forces, torques,
    Upot and virial of gpuLinkedCells/gpulc_c08 (newton3) and gpulc_c18 against LinkedCells/lc_c08 at 1e-12."""
    r = subprocess.run([BIN_MS], capture_output=True, text=True, timeout=600)
    print(r.stdout)
    print(r.stderr)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "SHIM TEST PASSED" in r.stdout


TUNER = os.path.join(ROOT, "oracle", "_ref", "autopas_tuner_driver")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(TUNER), reason="oracle/_ref/autopas_tuner_driver was not built (reference tree absent)")
def test_autotuner_samples_cpu_and_gpu_configurations_side_by_side():
    """A stock autopas::AutoPas<MoleculeLJ> (the reference's facade, LogicHandler and AutoTuner, compiled with the header
    overlay of tools/make_autopas_overlay.py = the additive edits of INTEGRATION.md) whose search space holds
    LinkedCells/lc_c08 next to gpuVerletClusterLists/gpuvcl_pruned and gpuLinkedCells/gpulc_c08, newton3 off and on:
    the tuner must sample all six through AutoPas::computeInteractions (container switches included), every sample must
    give the forces / Upot / virial of the CPU configuration, and the tuning phase must end with one of them chosen."""
    import json
    r = subprocess.run([TUNER, "24", "40"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    rep = json.loads(r.stdout)
    its = rep["iterations"]
    configs = {i["config"] for i in its}
    assert rep["num_configs_sampled"] == 6 and len(configs) == 6, configs
    assert {c.split("/")[0] for c in configs} == {"LinkedCells", "gpuVerletClusterLists", "gpuLinkedCells"}
    assert {c.split("/")[1] for c in configs} == {"lc_c08", "gpuvcl_pruned", "gpulc_c08"}
    for i in its:
        if i["max_rel_force_dev"] >= 0:
            assert i["max_rel_force_dev"] <= 1e-12, i
        assert i["upot"] == pytest.approx(rep["ref_upot"], rel=1e-12), i
        assert i["virial"] == pytest.approx(rep["ref_virial"], rel=1e-12), i
    assert its[0]["tuning"] and not its[-1]["tuning"]  # 6 configurations x 3 samples, then the optimum runs
    assert rep["chosen"] in configs and its[-1]["config"] == rep["chosen"]
    print("tuner chose", rep["chosen"])
