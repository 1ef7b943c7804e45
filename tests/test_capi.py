"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/autopas_b200.h
declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from autopas_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "autopas_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(apb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(declared) == set(capi.SIGNATURES), "ctypes binding and header disagree"


def test_struct_layouts_match_header():
    # field order / count follow the header; sizes are what a C compiler gives on x86-64
    assert ctypes.sizeof(capi.Config) == 6 * 8 + 3 * 8 + 4 * 4
    assert ctypes.sizeof(capi.TraversalResult) == 4 * 8 + 5 * 8
    assert ctypes.sizeof(capi.Geometry) == 3 * 8 + 3 * 8 + 8 + 6 * 8
    assert ctypes.sizeof(capi.Functor) == 8 + 3 * 8 + 8 + 8 + 8 + 8 + 3 * 8


def test_struct_layouts_match_the_c_compiler(tmp_path):
    """sizeof / offsetof of every ABI struct as gcc lays them out from include/autopas_b200.h vs the ctypes mirror."""
    import subprocess
    src = tmp_path / "layout.c"
    fields = {"apb_config": ["cutoff", "cluster_size", "device"],
              "apb_functor": ["cutoff", "num_types", "mixing_table", "nu", "num_mol_types", "site_start", "site_types"],
              "apb_traversal_result": ["virial_sum", "num_dist_calls", "num_global_calcs_no_n3"],
              "apb_geometry": ["cell_length", "num_slots", "towers_per_interaction_length"],
              "apb_loop_params": ["mass_of_type", "num_types", "global_force", "rebuild_frequency", "newton3"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "autopas_b200.h"', 'int main(void) {']
    for st, fs in fields.items():
        lines.append(f'printf("{st} %zu\\n", sizeof({st}));')
        for f in fs:
            lines.append(f'printf("{st}.{f} %zu\\n", offsetof({st}, {f}));')
    lines += ['return 0; }']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    mirror = {"apb_config": capi.Config, "apb_functor": capi.Functor, "apb_traversal_result": capi.TraversalResult,
              "apb_geometry": capi.Geometry, "apb_loop_params": capi.LoopParams}
    for st, fs in fields.items():
        assert int(out[st]) == ctypes.sizeof(mirror[st]), st
        for f in fs:
            assert int(out[f"{st}.{f}"]) == getattr(mirror[st], f).offset, f"{st}.{f}"


def test_host_helpers_match_reference_formulas():
    lib = capi.load()
    # ParticlePropertiesLibrary::calcShift6 (ParticlePropertiesLibrary.h:576-582)
    s2 = 1.0 / 6.25
    s6 = s2 * s2 * s2
    assert lib.apb_lj_calc_shift6(24.0, 1.0, 6.25) == 24.0 * (s6 - s6 * s6)
    raw = capi.TraversalResult()
    raw.upot_sum = 12.0
    raw.virial_sum[0], raw.virial_sum[1], raw.virial_sum[2] = 2.0, 4.0, 6.0
    raw.num_dist_calls, raw.num_kernel_calls_n3, raw.num_kernel_calls_no_n3 = 10, 3, 4
    raw.num_global_calcs_n3, raw.num_global_calcs_no_n3 = 3, 4
    u, v = ctypes.c_double(), ctypes.c_double()
    lib.apb_lj_end_traversal(ctypes.byref(raw), ctypes.byref(u), ctypes.byref(v))
    assert u.value == 12.0 * 0.5 / 6.0 and v.value == 6.0  # LJFunctor.h:661-685, :719
    assert lib.apb_lj_num_flops(ctypes.byref(raw), 1) == 10 * 8 + 3 * 18 + 4 * 15 + 3 * 13 + 4 * 9  # :776-789
    assert lib.apb_lj_num_flops(ctypes.byref(raw), 0) == 10 * 8 + 3 * 18 + 4 * 15 + 3 * 12 + 4 * 8


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from autopas_b200 import ApbError, GpuParticleContainer
    with pytest.raises(ApbError) as e:
        GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [10, 10, 10], 1.0, 0.2)
    assert e.value.code == capi.ERR_CUDA
    assert "no CPU fallback" in str(e.value)
    # without a handle there is nothing to compute with: the checkpoint entry points refuse as well (only the .pvtu
    # index, plain host text, needs no device)
    import ctypes
    lib, n = capi.load(), ctypes.c_int64()
    assert lib.apb_vtk_particle_record(None, None, 0, ctypes.byref(n)) == capi.ERR_INVALID_ARGUMENT
    assert lib.apb_vtk_write_particle_record(None, b"/tmp/never.vtu", ctypes.byref(n)) == capi.ERR_INVALID_ARGUMENT
    assert lib.apb_vtk_load_particle_record(None, b"<", 1, 1, ctypes.byref(n)) == capi.ERR_INVALID_ARGUMENT


def test_invalid_configs_are_rejected_before_touching_the_device():
    from autopas_b200 import ApbError, GpuParticleContainer
    with pytest.raises(ApbError) as e:
        GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [1, 1, 1], 1.0, 0.2)  # box < cutoff + skin
    assert e.value.code == capi.ERR_INVALID_ARGUMENT
    with pytest.raises(ApbError) as e:
        GpuParticleContainer("gpuVerletClusterLists", [0, 0, 0], [10, 10, 10], 1.0, 0.2, clusterSize=5)
    assert e.value.code == capi.ERR_NOT_APPLICABLE
    with pytest.raises(ApbError):
        GpuParticleContainer("linkedCells", [0, 0, 0], [10, 10, 10], 1.0, 0.2)


def test_lj_mixing_rules_match_the_reference_test():
    """ParticlePropertiesLibraryTest.cpp:254-342 (LennardJonesMixingTest) and :69-107 (AddingDifferentSitesTest) for the
    host-side table builder of the product library (apb_make_lj_mixing_table) and for the oracle: eps24_ij =
    24 sqrt(eps_i eps_j), sigma2_ij = ((sigma_i + sigma_j) / 2)^2, shift6_ij = calcShift6(eps24_ij, sigma2_ij, rc^2)."""
    import numpy as np

    import oracle
    from autopas_b200 import ParticlePropertiesLibrary

    cutoff = 1.1
    eps, sig = [0.6, 0.7, 1.0], [1.2, 1.4, 1.0]
    ppl = ParticlePropertiesLibrary(cutoff)
    for t in range(3):
        ppl.addSiteType(t, 1.0)
        ppl.addLJParametersToSite(t, eps[t], sig[t])
    ppl.calculateMixingCoefficients()
    table = oracle.mixing_table(eps, sig, cutoff).reshape(3, 3, 3)

    def shift6(e24, s2):
        s6 = (s2 / (cutoff * cutoff)) ** 3
        return e24 * (s6 - s6 * s6)

    for i in range(3):
        for j in range(3):
            e24 = 24.0 * np.sqrt(eps[i] * eps[j])
            s2 = ((sig[i] + sig[j]) / 2.0) ** 2
            for got in ((ppl.getMixing24Epsilon(i, j), ppl.getMixingSigmaSquared(i, j), ppl.getMixingShift6(i, j)), table[i, j]):
                assert got[0] == pytest.approx(e24, rel=4e-16)
                assert got[1] == pytest.approx(s2, rel=4e-16)
                assert got[2] == pytest.approx(shift6(e24, s2), rel=1e-14)
    # a pure LJ site next to a pure Axilrod-Teller site (zero-initialised parameters)
    mixed = ParticlePropertiesLibrary(0.1)
    mixed.addSiteType(0, 1.0)
    mixed.addLJParametersToSite(0, 1.0, 1.0)
    mixed.addSiteType(1, 1.2)
    mixed.addATMParametersToSite(1, 0.1)
    mixed.calculateMixingCoefficients()
    assert mixed.getMixing24Epsilon(0, 1) == 0.0 and mixed.getMixingSigmaSquared(1, 0) == 0.25
    nu = mixed.getMixingNuTable().reshape(2, 2, 2)
    assert nu[0, 0, 1] == 0.0 and nu[1, 1, 0] == 0.0 and nu[1, 1, 1] == pytest.approx(0.1, rel=1e-15)
