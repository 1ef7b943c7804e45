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
