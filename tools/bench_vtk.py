"""Timing of the device-side checkpoint record (apb_vtk_particle_record): bytes, ms for the size query (measuring pass +
scans) and for the full record into host memory, one JSON line per size."""
import ctypes
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autopas_b200 import GpuParticleContainer, capi

for n in [int(a) for a in sys.argv[1:]] or [1_000_000, 4_000_000]:
    rng = np.random.default_rng(3)
    L = (n / 0.8442) ** (1 / 3)
    c = GpuParticleContainer("gpuLinkedCells", [0, 0, 0], [L, L, L], 2.5, 0.3)
    c.addParticles(rng.uniform(0, L, n), rng.uniform(0, L, n), rng.uniform(0, L, n), np.arange(n, dtype=np.int64))
    for k in ("VX", "VY", "VZ", "FX", "FY", "FZ"):
        c.uploadColumn(k, rng.normal(size=n))
    lib, size = capi.load(), ctypes.c_int64()
    c.vtkParticleRecord()  # warm-up: buffers, tables
    t0 = time.perf_counter()
    lib.apb_vtk_particle_record(c._h, None, 0, ctypes.byref(size))
    t1 = time.perf_counter()
    buf = np.zeros(size.value, dtype=np.uint8)
    t2 = time.perf_counter()
    lib.apb_vtk_particle_record(c._h, buf.ctypes.data_as(ctypes.c_void_p), size.value, ctypes.byref(size))
    t3 = time.perf_counter()
    t4 = time.perf_counter()
    c.writeVtkParticleRecord("/tmp/apb_bench_vtk.vtu")
    t5 = time.perf_counter()
    os.remove("/tmp/apb_bench_vtk.vtu")
    print(json.dumps({"particles": n, "ms_record_to_file": (t5 - t4) * 1e3, "record_bytes": size.value, "ms_size_query": (t1 - t0) * 1e3, "ms_record_to_host": (t3 - t2) * 1e3,
                      "GB_per_s_text": size.value / (t3 - t2) / 1e9}))
    c.close()
