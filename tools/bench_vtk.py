"""Timing of the device-side checkpoint record (apb_vtk_particle_record / apb_vtk_write_particle_record): bytes, ms for the
size query (measuring pass + scans), for the record into host memory and into a file; one JSON line per size."""
import ctypes
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autopas_b200 import GpuParticleContainer, capi  # noqa: E402


def record_timing(n):
    """A liquid-density box of n particles with normally distributed velocities and forces."""
    rng = np.random.default_rng(3)
    # (a whole-number box edge: writeWithDynamicPrecision only protects positions that round exactly onto the border)
    L = float(np.ceil((n / 0.8442) ** (1 / 3)))
    c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], 2.5, 0.3)
    try:
        c.addParticles(rng.uniform(0, L, n), rng.uniform(0, L, n), rng.uniform(0, L, n), np.arange(n, dtype=np.int64))
        for k in ("VX", "VY", "VZ", "FX", "FY", "FZ"):
            c.uploadColumn(k, rng.normal(size=n))
        lib, size = capi.load(), ctypes.c_int64()
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "bench_Particles_0_000000.vtu")
            c.writeVtkParticleRecord(path)  # warm-up: buffers, tables, pinned staging
            t0 = time.perf_counter()
            lib.apb_vtk_particle_record(c._h, None, 0, ctypes.byref(size))
            t1 = time.perf_counter()
            buf = np.zeros(size.value, dtype=np.uint8)
            t2 = time.perf_counter()
            lib.apb_vtk_particle_record(c._h, buf.ctypes.data_as(ctypes.c_void_p), size.value, ctypes.byref(size))
            t3 = time.perf_counter()
            c.writeVtkParticleRecord(path)
            t4 = time.perf_counter()
        d = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], 2.5, 0.3)
        try:
            t5 = time.perf_counter()
            loaded = d.loadVtkParticleRecord(buf)
            t6 = time.perf_counter()
        finally:
            d.close()
        assert loaded == n
        return {"config": "md-flexible VTK checkpoint (ParallelVtkWriter::recordParticleStates) formatted on the device, and read back (loadParticlesFromRankRecord)",
                "ms_load_from_host": (t6 - t5) * 1e3,
                "particles": n, "record_bytes": size.value, "ms_size_query": (t1 - t0) * 1e3, "ms_record_to_host": (t3 - t2) * 1e3,
                "ms_record_to_file": (t4 - t3) * 1e3, "GB_per_s_text_to_file": size.value / (t4 - t3) / 1e9}
    finally:
        c.close()


if __name__ == "__main__":
    for n_ in [int(a) for a in sys.argv[1:]] or [1_000_000, 4_000_000]:
        print(json.dumps(record_timing(n_)))
